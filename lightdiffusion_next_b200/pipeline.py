"""pipe(prompt, ...) / sample() / decode() surface of the engine — the host-side mirror of `pipeline()`
(src/user/pipeline.py:31-518) for the SD1.5 txt2img path: CLIP encode -> KSampler -> VAE decode.

Tokenisation stays with the caller (the reference's SD1Tokenizer, src/SD15/SDToken.py:292-396, is host Python that runs in
microseconds and needs its vocabulary files): `tokens` is what `tokenize_with_weights` returns, i.e. a list of 77-long
lists of (token_id, weight).  Everything after that runs on the B200 engine.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import sampling as S
from .engine import Engine

EMPTY_TOKENS = [49406] + [49407] * 76  # <|startoftext|>, <|endoftext|> padding: what the SD1 tokenizer emits for ""


PAD_TOKEN = 49407


def resolve_textual_embeddings(tokens, vocab: int, width: int = 768, pad: int = PAD_TOKEN):
    """SDClipModel.set_up_textual_embeddings (src/SD15/SDClip.py:213-268): a token row may carry textual-inversion embedding
    VECTORS in place of ids (what the reference's tokenizer emits for "embedding:name"; its default negative prompt has four,
    src/user/pipeline.py:98).  Every vector of the right width gets the next id past the vocabulary; a vector of the wrong
    width is dropped (the ids behind it move up and the row is re-padded at its end -- the weights keep their positions,
    as in the reference).  Returns (ids [rows, n] int64, weights [rows, n] fp32, [vectors in id order])."""
    extra: List[torch.Tensor] = []
    ids_rows, w_rows = [], []
    for row in tokens:
        ids = []
        for t, _ in row:
            if isinstance(t, torch.Tensor):
                if t.dim() == 1 and t.shape[0] == width:
                    ids.append(vocab + len(extra))
                    extra.append(t)
                # else: ignored, like the reference (it logs a warning)
            else:
                if int(t) < 0:
                    raise ValueError("negative token ids are not supported")
                ids.append(int(t))
        ids += [pad] * (len(row) - len(ids))
        ids_rows.append(ids)
        w_rows.append([float(w) for _, w in row])
    return torch.tensor(ids_rows, dtype=torch.int64), torch.tensor(w_rows, dtype=torch.float32), extra


def extend_token_table(base: torch.Tensor, vectors: Sequence[torch.Tensor]) -> torch.Tensor:
    """The reference builds a new Embedding in the table's own dtype with the vectors appended (SDClip.py:247-259): the
    vectors are rounded to that dtype (fp16 checkpoints) exactly as there."""
    return torch.cat([base.cpu(), torch.stack([v.detach().cpu() for v in vectors]).to(base.dtype)])


class Pipeline:
    def __init__(self, engine: Engine):
        self.e = engine

    # ---------------------------------------------------------------- CLIPTextEncode (src/clip/Clip.py:574-589)
    def encode(self, tokens: Sequence[Sequence[Tuple[int, float]]], return_pooled: bool = False):
        """tokens: k chunks of 77 (id, weight) pairs -> conditioning [1, 77*k, 768] (layer -2 + final LN).
        Prompt weights: per-token lerp against the empty-prompt encoding (ClipTokenWeightEncoder, SDClip.py:54-76).
        return_pooled: also return the pooled vector [1, 768] of the first chunk -- the last layer's final-LN state at the
        first end-of-text token (CLIPTextModel_.forward, src/clip/CLIPTextModel.py:95-105; the `y` input of Flux)."""
        if any(isinstance(t, torch.Tensor) for row in tokens for t, _ in row):
            ids, wts, extra = resolve_textual_embeddings(tokens, self.e.clip_vocab())
            self.e.set_clip_extra_embeddings(extra)
        else:
            ids = torch.tensor([[t for t, _ in row] for row in tokens], dtype=torch.int64)
            wts = torch.tensor([[w for _, w in row] for row in tokens], dtype=torch.float32)
        has_w = bool((wts != 1.0).any())
        if has_w:
            ids = torch.cat([ids, torch.tensor([EMPTY_TOKENS], dtype=torch.int64)])
        pen, last = self.e.clip_encode(ids)
        if return_pooled:
            eos = int((ids[0] == EMPTY_TOKENS[-1]).int().argmax())
            pooled = last[0:1, eos].clone()
        if has_w:
            z_empty = pen[-1]
            pen = pen[:-1]
            w = wts.to(pen.device)[:, :, None]
            pen = (pen - z_empty) * w + z_empty
        cond = pen.reshape(1, -1, pen.shape[-1])
        return (cond, pooled) if return_pooled else cond

    # ---------------------------------------------------------------- KSampler (src/sample/sampling.py:773-887)
    def sample(self, positive: torch.Tensor, negative: torch.Tensor, width: int, height: int, batch: int = 1,
               seed: int = 0, steps: int = 20, cfg: float = 7.0, sampler_name: str = "dpmpp_2m_cfgpp",
               scheduler: str = "karras", enable_multiscale: bool = True,
               sampler_options: Optional[Dict[str, object]] = None) -> torch.Tensor:
        latent = {"samples": torch.zeros(batch, 4, height // 8, width // 8)}  # EmptyLatentImage (Latent.py:174-190)
        return S.sample(self.e, seed, steps, cfg, sampler_name, scheduler, positive, negative, latent,
                        enable_multiscale=enable_multiscale, sampler_options=sampler_options)[0]["samples"]

    # ---------------------------------------------------------------- VAEDecode (VariationalAE.py:771-784)
    def decode(self, samples: torch.Tensor) -> torch.Tensor:
        """latents (already divided by 0.18215, as KSampler returns them) -> [B,H,W,3] fp32 in [0,1] on the CPU."""
        return self.e.vae_decode(samples).cpu()

    # ---------------------------------------------------------------- VAEEncode + img2img (VariationalAE.py:787-801)
    def encode_image(self, pixels: torch.Tensor) -> torch.Tensor:
        """pixels [B,H,W,3] in [0,1] -> latent samples as VAEEncode returns them (unscaled; KSampler applies 0.18215)."""
        return self.e.vae_encode(pixels)

    def img2img(self, pixels: torch.Tensor, positive: torch.Tensor, negative: torch.Tensor, seed: int = 0, steps: int = 20,
                cfg: float = 7.0, denoise: float = 0.6, sampler_name: str = "dpmpp_2m_cfgpp", scheduler: str = "karras",
                enable_multiscale: bool = True) -> torch.Tensor:
        """Encode -> partial-denoise KSampler pass (the schedule tail of steps/denoise, sampling.py:655-675) -> latents."""
        latent = {"samples": self.encode_image(pixels)}
        return S.sample(self.e, seed, steps, cfg, sampler_name, scheduler, positive, negative, latent, denoise=denoise,
                        enable_multiscale=enable_multiscale)[0]["samples"]

    # ---------------------------------------------------------------- taesd_preview (taesd.py:219-255)
    def preview(self, x: torch.Tensor) -> torch.Tensor:
        """Current sampler latent [B,4,h,w] -> preview images [B,8h,8w,3] in [0,1] on the CPU (what the reference shows every
        5 steps); usable as / from a sampler `callback`."""
        return self.e.taesd_decode(x[:, :4]).clamp_(0.0, 1.0).cpu()

    def __call__(self, tokens, negative_tokens=None, width: int = 512, height: int = 512, batch: int = 1, seed: int = 0,
                 steps: int = 20, cfg: float = 7.0, sampler_name: str = "dpmpp_2m_cfgpp", scheduler: str = "karras"):
        pos = self.encode(tokens)
        neg = self.encode(negative_tokens if negative_tokens is not None else [[(t, 1.0) for t in EMPTY_TOKENS]])
        lat = self.sample(pos, neg, width, height, batch, seed, steps, cfg, sampler_name, scheduler)
        return self.decode(lat)


def hires_fix(pipe: "Pipeline", samples: torch.Tensor, positive: torch.Tensor, negative: torch.Tensor, width: int,
              height: int, seed: int = 0, steps: int = 10, cfg: float = 8.0, denoise: float = 0.45,
              sampler_name: str = "euler_ancestral_cfgpp", scheduler: str = "normal") -> torch.Tensor:
    """The HiresFix branch of pipeline() (src/user/pipeline.py:346-366): bislerp-upscale the latent to 2x the pixel size,
    then a second KSampler pass with partial denoise.  `samples`: latents as returned by Pipeline.sample."""
    from .latent import latent_upscale

    up = latent_upscale({"samples": samples}, width * 2, height * 2, engine=pipe.e if hasattr(pipe.e, "bislerp") else None)
    return S.sample(pipe.e, seed, steps, cfg, sampler_name, scheduler, positive, negative, up, denoise=denoise)[0]["samples"]


_hires_fix_pass = hires_fix  # pipeline() below has a flag of the same name


# the reference's default negative prompt (src/user/pipeline.py:98); its four "embedding:" terms are textual-inversion
# files the tokenizer resolves from its embedding directory (absent files are skipped with a warning, SDToken.py:335-352)
DEFAULT_NEGATIVE_PROMPT = ("(worst quality, low quality:1.4), (zombie, sketch, interlocked fingers, comic), "
                           "(embedding:EasyNegative), (embedding:badhandv4), (embedding:lr), (embedding:ng_deepnegative_v1_75t)")
last_seed = 0  # module state like the reference's `last_seed` (pipeline.py:22, 102-107)


def _token_rows(tokenizer, text: str):
    """SD1Tokenizer.tokenize_with_weights returns {"l": rows}, SDTokenizer returns the rows (SDToken.py:292-396, 425-440)."""
    t = tokenizer.tokenize_with_weights(text)
    return t["l"] if isinstance(t, dict) else t


def pipeline(engine: Engine, prompt: str, w: int, h: int, number: int = 1, batch: int = 1, hires_fix: bool = False,
             adetailer: bool = False, enhance_prompt: bool = False, img2img: bool = False, stable_fast: bool = False,
             reuse_seed: bool = False, flux_enabled: bool = False, prio_speed: bool = False, autohdr: bool = False,
             realistic_model: bool = False, negative_prompt: Optional[str] = None, multiscale_preset: Optional[str] = None,
             enable_multiscale: bool = True, multiscale_factor: float = 0.5, multiscale_fullres_start: int = 3,
             multiscale_fullres_end: int = 8, multiscale_intermittent_fullres: bool = False, *, tokenizer=None,
             seed: Optional[int] = None, hires_seed: Optional[int] = None) -> List[torch.Tensor]:
    """`pipeline(prompt, w, h, number, batch, ...)` of the reference (src/user/pipeline.py:31-518), SD1.5 txt2img branch
    (:278-372), on an engine that already holds the UNet / VAE / CLIP weights (checkpoint + LoRA loading is
    `Engine.load_checkpoint`): CLIPSetLastLayer(-2) + CLIPTextEncode of prompt and negative prompt -> EmptyLatentImage ->
    KSampler(20 steps, cfg 7, dpmpp_sde_cfgpp -- dpmpp_2m_cfgpp with prio_speed --, karras, the caller's multiscale options)
    -> [HiresFix: LatentUpscale x2 + KSampler(10 steps, cfg 8, euler_ancestral_cfgpp, normal, denoise 0.45)] -> VAEDecode.

    Same positional / keyword arguments as the reference.  Differences, all at the edges of the hot path: the images are
    RETURNED (a list of `number` tensors [batch, H, W, 3] in [0, 1]) instead of being written as PNGs; `tokenizer` is the
    reference's tokenizer object (SD1Tokenizer / SDTokenizer: host Python + vocabulary files, out of scope) or anything with
    its `tokenize_with_weights(text)`; `seed` pins the seed the reference draws at random (`reuse_seed` works as there).
    The flags for subsystems outside SURVEY.md 8 raise instead of being ignored: adetailer, enhance_prompt, autohdr,
    img2img-from-a-path (use Pipeline.img2img with pixels), flux_enabled (use FluxPipeline).  `stable_fast` and
    `realistic_model` are accepted and inert: the engine replaces the former, the latter picks a checkpoint file.
    As in the reference, `enable_multiscale` & co. only reach dpmpp_sde_cfgpp; dpmpp_2m_cfgpp runs its own defaults (fact 9)."""
    import random

    global last_seed
    for flag, name in ((adetailer, "adetailer"), (enhance_prompt, "enhance_prompt"), (autohdr, "autohdr"),
                       (img2img, "img2img"), (flux_enabled, "flux_enabled")):
        if flag:
            raise NotImplementedError(f"pipeline({name}=True) is outside the sampler hot path this engine replaces")
    if tokenizer is None:
        raise ValueError("pipeline() needs tokenizer= (the reference's SD1Tokenizer, src/SD15/SDToken.py:399-440)")
    if multiscale_preset is not None:
        enable_multiscale, multiscale_factor, multiscale_fullres_start, multiscale_fullres_end, \
            multiscale_intermittent_fullres = MULTISCALE_PRESETS[multiscale_preset]
    if negative_prompt is None or negative_prompt.strip() == "":
        negative_prompt = DEFAULT_NEGATIVE_PROMPT
    if reuse_seed:
        seed = last_seed
    elif seed is None:
        seed = random.randint(1, 2 ** 64)
    last_seed = seed
    sampler_name = "dpmpp_sde_cfgpp" if not prio_speed else "dpmpp_2m_cfgpp"
    pipe = Pipeline(engine)
    images = []
    for _ in range(number):
        pos = pipe.encode(_token_rows(tokenizer, prompt))
        neg = pipe.encode(_token_rows(tokenizer, negative_prompt))
        opts = None
        if sampler_name == "dpmpp_sde_cfgpp":  # the only sampler sample1 forwards these to (sampling.py:949-964)
            opts = {"multiscale_factor": multiscale_factor, "multiscale_fullres_start": multiscale_fullres_start,
                    "multiscale_fullres_end": multiscale_fullres_end,
                    "multiscale_intermittent_fullres": multiscale_intermittent_fullres}
        lat = pipe.sample(pos, neg, w, h, batch, seed, 20, 7.0, sampler_name, "karras", enable_multiscale=enable_multiscale,
                          sampler_options=opts)
        if hires_fix:
            hs = hires_seed if hires_seed is not None else random.randint(1, 2 ** 64)
            lat = _hires_fix_pass(pipe, lat, pos, neg, w, h, seed=hs)
        images.append(pipe.decode(lat))
    return images


# src/sample/multiscale_presets.py:49-86: (enable, factor, fullres_start, fullres_end, intermittent)
MULTISCALE_PRESETS = {"quality": (True, 0.5, 10, 8, True), "performance": (True, 0.25, 5, 8, True),
                      "balanced": (True, 0.5, 5, 8, True), "disabled": (False, 1.0, 0, 0, False)}


class FluxPipeline:
    """The Flux branch of `pipeline()` (src/user/pipeline.py:218-275) on one engine holding the Flux DiT (load_flux), the
    T5-XXL and CLIP-L text encoders (load_t5 / load_clip) and the 16-channel autoencoder (load_vae):
    CLIPTextEncodeFlux -> ConditioningZeroOut -> KSampler(euler_cfgpp, beta, cfg 1, flux=True) -> VAEDecode(flux=True).
    Tokenisation stays with the caller: `clip_tokens` as for `Pipeline.encode`, `t5_tokens` = the T5 tokenizer's rows of
    (id, weight) (t5.pad_tokens builds one from sentencepiece ids)."""

    def __init__(self, engine: Engine):
        self.e = engine
        self._clip = Pipeline(engine)

    def encode(self, clip_tokens, t5_tokens, guidance: float = 3.0) -> Dict[str, object]:
        """CLIPTextEncodeFlux.encode (src/Quantize/Quantizer.py:960-990): FluxClipModel.encode_token_weights returns the T5
        states as the conditioning and CLIP-L's pooled vector (src/clip/FluxClip.py:704-718)."""
        from . import t5 as T5H

        _, pooled = self._clip.encode(clip_tokens, return_pooled=True)
        return {"cond": T5H.encode_token_weights(self.e, t5_tokens), "pooled_output": pooled, "guidance": float(guidance)}

    @staticmethod
    def zero_out(conditioning: Dict[str, object]) -> Dict[str, object]:
        """ConditioningZeroOut.zero_out (src/Quantize/Quantizer.py:993-1012): the negative of the Flux branch."""
        return {"cond": torch.zeros_like(conditioning["cond"]), "pooled_output": torch.zeros_like(conditioning["pooled_output"]),
                "guidance": conditioning["guidance"]}

    def sample(self, positive: Dict[str, object], negative: Dict[str, object], width: int, height: int, batch: int = 1,
               seed: int = 0, steps: int = 20, cfg: float = 1.0, scheduler: str = "beta") -> torch.Tensor:
        """EmptyLatentImage (4 zero channels, repeated to the model's 16 by fix_empty_latent_channels,
        src/Utilities/Latent.py:192-204) -> the Flux KSampler call of pipeline.py:251-264."""
        from . import flux_sampling as FS

        latent = {"samples": torch.zeros(batch, 16, height // 8, width // 8)}
        return FS.sample_flux(self.e, seed, steps, (positive["cond"], positive["pooled_output"]),
                              (negative["cond"], negative["pooled_output"]), latent, cfg=cfg,
                              guidance=float(positive["guidance"]), scheduler=scheduler)[0]["samples"]

    def decode(self, samples: torch.Tensor) -> torch.Tensor:
        """VAEDecode(flux=True): latents in the autoencoder's space (process_out applied by the sampler) -> [B,H,W,3] in [0,1]."""
        return self.e.vae_decode(samples).cpu()

    def __call__(self, clip_tokens, t5_tokens, width: int = 1024, height: int = 1024, batch: int = 1, seed: int = 0,
                 steps: int = 20, guidance: float = 3.0) -> torch.Tensor:
        pos = self.encode(clip_tokens, t5_tokens, guidance)
        return self.decode(self.sample(pos, self.zero_out(pos), width, height, batch, seed, steps))

