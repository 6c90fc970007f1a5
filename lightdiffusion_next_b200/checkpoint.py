"""Weight ingest from real SD1.5 checkpoints: file -> the three state dicts the engine loads.

Mirrors the reference's on-disk contract, not its loader classes:
  * container: `.safetensors` (or a pickled `.ckpt` / `.pt` with an optional top-level "state_dict"), as
    `load_torch_file` accepts (src/Utilities/util.py, used from src/FileManaging/Loader.py:11-111);
  * key prefixes of an LDM SD1.5 checkpoint: `model.diffusion_model.` (UNet), `first_stage_model.` (VAE),
    `cond_stage_model.transformer.[text_model.]` (CLIP-L; the older layout without `text_model.` is renamed exactly as
    SD15.process_clip_state_dict does, src/SD15/SD15.py:32-69);
  * LoRA merge: kohya-style `lora_unet_*` / `lora_te_text_model_encoder_layers_*` keys (src/Model/LoRas.py:15-121) folded
    into the weights with `W += strength * (alpha / rank) * (up @ down)` (ModelPatcher.calculate_weight,
    src/Model/ModelPatcher.py:621-650, applied by patch_weight_to_device :267-300): fp32 product added to an fp32 copy of
    the weight, rounded once to the weight's storage dtype.

The engine never sees file formats: it takes the resulting `{name: tensor}` dicts through `Engine.load_unet / load_vae /
load_clip` (C ABI `ldn_load_weights`), which repack to the HBM layout once.
"""
from __future__ import annotations

from typing import Dict, Iterable, Mapping, Optional, Tuple

import torch

from . import synth

UNET_PREFIX = "model.diffusion_model."
VAE_PREFIX = "first_stage_model."
CLIP_PREFIXES = ("cond_stage_model.transformer.text_model.", "cond_stage_model.transformer.")

# src/Model/LoRas.py:5-12
_LORA_CLIP_MAP = {
    "mlp.fc1": "mlp_fc1",
    "mlp.fc2": "mlp_fc2",
    "self_attn.k_proj": "self_attn_k_proj",
    "self_attn.q_proj": "self_attn_q_proj",
    "self_attn.v_proj": "self_attn_v_proj",
    "self_attn.out_proj": "self_attn_out_proj",
}


def load_state_dict_file(path: str) -> Dict[str, torch.Tensor]:
    """Read a checkpoint container into a flat `{key: cpu tensor}` dict."""
    if path.endswith(".gguf"):
        return load_gguf(path, handle_prefix=None)
    if path.endswith((".safetensors", ".sft")):
        from safetensors.torch import load_file

        return load_file(path, device="cpu")
    sd = torch.load(path, map_location="cpu", weights_only=True)
    if isinstance(sd, dict) and "state_dict" in sd and isinstance(sd["state_dict"], dict):
        sd = sd["state_dict"]
    return sd


def load_gguf(path: str, handle_prefix: Optional[str] = "model.diffusion_model.", dtype: torch.dtype = torch.bfloat16) -> Dict[str, torch.Tensor]:
    """GGUF container (Flux / SD checkpoints quantised by llama.cpp-style tools) -> plain `{key: cpu tensor}`.

    The reference keeps Q8_0 weights packed and dequantises them inside every Linear call (GGMLOps,
    src/Quantize/Quantizer.py:352-390; gguf_sd_loader :581-666; dequantize_blocks_Q8_0 :94-104). Here the blocks are
    dequantised ONCE at ingest -- w = d * q with d the block's fp16 scale and q its 32 int8 values, exactly the reference's
    arithmetic -- and stored as bf16 in HBM (a B200 holds the 24 GB of Flux.1-dev in bf16 eight times over). F32 / F16 / BF16
    tensors pass through; other quantisation types are rejected, as the reference's dequantize table only knows Q8_0.
    Uses the same third-party `gguf` reader the reference depends on (requirements: gguf)."""
    import gguf
    import numpy as np

    reader = gguf.GGUFReader(path)
    arch = reader.get_field("general.architecture")
    if arch is not None:  # same check as gguf_sd_loader (:600-614)
        if len(arch.types) != 1 or arch.types[0] != gguf.GGUFValueType.STRING:
            raise TypeError(f"{path}: bad type for GGUF general.architecture: expected string, got {arch.types!r}")
        arch_str = str(bytes(arch.parts[arch.data[-1]]), encoding="utf-8")
        if arch_str not in {"flux", "sd1", "sdxl", "t5", "t5encoder"}:
            raise ValueError(f"{path}: unexpected GGUF architecture {arch_str!r} (flux, sd1, sdxl, t5, t5encoder are accepted)")
    names = [t.name for t in reader.tensors]
    strip = len(handle_prefix) if handle_prefix and any(n.startswith(handle_prefix) for n in names) else 0
    out: Dict[str, torch.Tensor] = {}
    Q = gguf.GGMLQuantizationType
    for t in reader.tensors:
        if strip and not t.name.startswith(handle_prefix):
            continue
        key = t.name[strip:]
        shape = _gguf_orig_shape(reader, t.name)  # converters that reshape a tensor for quantisation record the original
        if shape is None:
            shape = tuple(int(v) for v in reversed(t.shape))  # GGUF stores dimensions innermost-first
        data = np.asarray(t.data)
        if t.tensor_type == Q.F32:
            w = torch.from_numpy(data.copy()).view(torch.float32).reshape(shape)
        elif t.tensor_type == Q.F16:
            w = torch.from_numpy(data.copy()).view(torch.float16).reshape(shape)
        elif t.tensor_type == Q.BF16:
            w = torch.from_numpy(data.copy()).view(torch.bfloat16).reshape(shape)
        elif t.tensor_type == Q.Q8_0:
            blocks = torch.from_numpy(data.copy()).view(torch.uint8).reshape(-1, 34)  # 2-byte fp16 scale + 32 int8
            d = blocks[:, :2].contiguous().view(torch.float16).to(torch.float32)
            q = blocks[:, 2:].contiguous().view(torch.int8).to(torch.float32)
            w = (d * q).reshape(shape).to(dtype)
        else:
            raise ValueError(f"GGUF tensor {t.name!r} has unsupported type {t.tensor_type!r} (F32, F16, BF16, Q8_0 are handled)")
        out[key] = w
    return out


def _gguf_orig_shape(reader, tensor_name: str) -> Optional[Tuple[int, ...]]:
    """`comfy.gguf.orig_shape.<tensor>` metadata (ARRAY of INT32), as gguf_sd_loader_get_orig_shape reads it (:427-447)."""
    import gguf

    key = f"comfy.gguf.orig_shape.{tensor_name}"
    field = reader.get_field(key)
    if field is None:
        return None
    if len(field.types) != 2 or field.types[0] != gguf.GGUFValueType.ARRAY or field.types[1] != gguf.GGUFValueType.INT32:
        raise TypeError(f"bad original shape metadata for {key}: expected ARRAY of INT32, got {field.types}")
    return tuple(int(field.parts[i][0]) for i in field.data)


# llama.cpp tensor names of a T5 encoder -> state-dict keys of the reference's T5 module, applied as successive substring
# replacements in this order (clip_sd_map / gguf_clip_loader, src/Quantize/Quantizer.py:815-858).
T5_GGUF_KEY_MAP = (
    ("enc.", "encoder."), (".blk.", ".block."), ("token_embd", "shared"), ("output_norm", "final_layer_norm"),
    ("attn_q", "layer.0.SelfAttention.q"), ("attn_k", "layer.0.SelfAttention.k"), ("attn_v", "layer.0.SelfAttention.v"),
    ("attn_o", "layer.0.SelfAttention.o"), ("attn_norm", "layer.0.layer_norm"),
    ("attn_rel_b", "layer.0.SelfAttention.relative_attention_bias"), ("ffn_up", "layer.1.DenseReluDense.wi_1"),
    ("ffn_down", "layer.1.DenseReluDense.wo"), ("ffn_gate", "layer.1.DenseReluDense.wi_0"), ("ffn_norm", "layer.1.layer_norm"),
)


def load_t5_gguf(path: str, dtype: torch.dtype = torch.bfloat16) -> Dict[str, torch.Tensor]:
    """`t5-v1_1-xxl-encoder-Q8_0.gguf` (the Flux text encoder the reference's pipeline loads, src/user/pipeline.py:233-237)
    -> the T5 state dict `Engine.load_t5` takes.  Like the reference, a file without encoder feed-forward tensors is refused."""
    raw = load_gguf(path, handle_prefix=None, dtype=dtype)
    if not any(k.startswith("enc.blk.") and k.endswith(".ffn_up.weight") for k in raw):
        raise ValueError(f"{path}: not a T5 encoder GGUF (no enc.blk.N.ffn_up.weight tensors)")
    out: Dict[str, torch.Tensor] = {}
    for k, v in raw.items():
        for a, b in T5_GGUF_KEY_MAP:
            k = k.replace(a, b)
        out[k] = v
    return out


def _strip(sd: Mapping[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def split_sd15_checkpoint(sd: Mapping[str, torch.Tensor], strict: bool = True) -> Dict[str, Dict[str, torch.Tensor]]:
    """Full-checkpoint state dict -> {"unet": ..., "vae": ..., "clip": ...} in the engine's key layout
    (UNet keys below `model.diffusion_model.`, VAE decoder-side keys below `first_stage_model.`, CLIP keys below
    `text_model.`).  Parts that are absent from the file are returned empty; with `strict` every present part is checked
    against the SD1.5 layout tables (names and shapes) and a ValueError lists what is missing or mis-shaped."""
    unet = _strip(sd, UNET_PREFIX)
    vae_all = _strip(sd, VAE_PREFIX)
    vae = {k: v for k, v in vae_all.items() if k.startswith(("decoder.", "post_quant_conv.", "encoder.", "quant_conv."))}
    clip: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        if k.startswith(CLIP_PREFIXES[0]):
            clip[k[len(CLIP_PREFIXES[0]):]] = v
        elif k.startswith(CLIP_PREFIXES[1]):  # pre-`text_model.` layout (SD15.py:43-53)
            clip[k[len(CLIP_PREFIXES[1]):]] = v
    clip.pop("embeddings.position_ids", None)  # a buffer, not a weight
    parts = {"unet": unet, "vae": vae, "clip": clip}
    if strict:
        problems = []
        vae_table = dict(synth.vae_decoder_shapes())
        if any(k.startswith("encoder.") for k in vae):  # encoder side is optional (VAE-decode-only files exist)
            vae_table.update(synth.vae_encoder_shapes())
        for name, table in (("unet", synth.unet_shapes()), ("vae", vae_table), ("clip", synth.clip_shapes())):
            part = parts[name]
            if not part:
                continue
            for k, shape in table.items():
                if k not in part:
                    problems.append(f"{name}: missing {k}")
                elif tuple(part[k].shape) != tuple(shape):
                    # 1x1 convs stored as linear weights (or vice versa) are the same tensor
                    if part[k].numel() == _numel(shape) and _squeeze(part[k].shape) == _squeeze(shape):
                        part[k] = part[k].reshape(shape)
                    else:
                        problems.append(f"{name}: {k} has shape {tuple(part[k].shape)}, expected {tuple(shape)}")
            parts[name] = {k: part[k] for k in table if k in part}  # drop keys the SD1.5 path does not use
        if problems:
            raise ValueError("not an SD1.5 checkpoint this engine can load:\n  " + "\n  ".join(problems[:20])
                             + (f"\n  ... and {len(problems) - 20} more" if len(problems) > 20 else ""))
    return parts


def _numel(shape: Iterable[int]) -> int:
    n = 1
    for s in shape:
        n *= int(s)
    return n


def _squeeze(shape: Iterable[int]) -> Tuple[int, ...]:
    return tuple(int(s) for s in shape if int(s) != 1)


_DIFFUSERS_RESNET = {"in_layers.0": "norm1", "in_layers.2": "conv1", "emb_layers.1": "time_emb_proj", "out_layers.0": "norm2",
                     "out_layers.3": "conv2", "skip_connection": "conv_shortcut"}
_DIFFUSERS_BASIC = {"input_blocks.0.0": "conv_in", "out.0": "conv_norm_out", "out.2": "conv_out",
                    "time_embed.0": "time_embedding.linear_1", "time_embed.2": "time_embedding.linear_2"}


def diffusers_unet_name(key: str, num_res: int = 2, levels: int = 4) -> Optional[str]:
    """LDM UNet weight key -> the name the same tensor has in a diffusers `UNet2DConditionModel` (the inverse direction of
    the reference's unet_to_diffusers table, src/NeuralNetwork/unet.py:85-185), for the SD1.x block layout: every level has
    `num_res` ResBlocks, a down/upsampler closes each level but the last / first.  None for keys without a counterpart."""
    stem, _, leaf = key.rpartition(".")
    if stem in _DIFFUSERS_BASIC:
        return f"{_DIFFUSERS_BASIC[stem]}.{leaf}"
    parts = key.split(".")
    per = num_res + 1

    def inner(sub: str, prefix: str, res_i, att_i) -> Optional[str]:
        # sub: "0.<resnet member>", "1.<transformer member>", "<j>.op.*" / "<j>.conv.*" handled by the callers
        j, _, rest = sub.partition(".")
        if j == "0":
            mod, _, lf = rest.rpartition(".")
            return f"{prefix}.resnets.{res_i}.{_DIFFUSERS_RESNET[mod]}.{lf}" if mod in _DIFFUSERS_RESNET else None
        if j == "1" and not rest.startswith("conv."):
            return f"{prefix}.attentions.{att_i}.{rest}"
        return None

    if parts[0] == "input_blocks":
        n = int(parts[1])
        x, i = divmod(n - 1, per)
        sub = ".".join(parts[2:])
        if i == num_res:  # the level's downsampler
            return f"down_blocks.{x}.downsamplers.0.conv.{leaf}" if sub.startswith("0.op.") else None
        return inner(sub, f"down_blocks.{x}", i, i)
    if parts[0] == "middle_block":
        j, sub = parts[1], ".".join(parts[2:])
        if j == "1":
            return f"mid_block.attentions.0.{sub}"
        mod, _, lf = sub.rpartition(".")
        return f"mid_block.resnets.{0 if j == '0' else 1}.{_DIFFUSERS_RESNET[mod]}.{lf}" if mod in _DIFFUSERS_RESNET else None
    if parts[0] == "output_blocks":
        n = int(parts[1])
        x, i = divmod(n, per)
        sub = ".".join(parts[2:])
        if len(parts) > 3 and parts[3] == "conv":  # the level's upsampler sits after the ResBlock (and the transformer)
            return f"up_blocks.{x}.upsamplers.0.conv.{leaf}"
        return inner(sub, f"up_blocks.{x}", i, i)
    return None


def lora_key_map(parts: Mapping[str, Mapping[str, torch.Tensor]]) -> Dict[str, Tuple[str, str]]:
    """LoRA module name -> (part, weight key).  UNet (model_lora_keys_unet, LoRas.py:86-121): `lora_unet_` / `lora_prior_unet_`
    + the LDM key with dots replaced by underscores, the same with the tensor's diffusers name (what kohya-trained SD1.5
    LoRAs use: `lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn1_to_q`), and the diffusers-LoRA spellings
    (`[unet.]<diffusers name>` with `.to_` -> `.processor.to_`, `to_out.0` -> `to_out`); CLIP: the three text-encoder
    spellings of LoRas.py:69-83."""
    m: Dict[str, Tuple[str, str]] = {}
    for k in parts.get("unet", {}):
        if k.endswith(".weight"):
            stem = k[: -len(".weight")]
            m["lora_unet_" + stem.replace(".", "_")] = ("unet", k)
            m["lora_prior_unet_" + stem.replace(".", "_")] = ("unet", k)
            d = diffusers_unet_name(k)
            if d is not None:
                dstem = d[: -len(".weight")]
                m["lora_unet_" + dstem.replace(".", "_")] = ("unet", k)
                proc = dstem.replace(".to_", ".processor.to_")
                if proc.endswith(".to_out.0"):
                    proc = proc[:-2]
                m[proc] = ("unet", k)
                m["unet." + proc] = ("unet", k)
    for k in parts.get("clip", {}):
        if not (k.startswith("encoder.layers.") and k.endswith(".weight")):
            continue
        _, _, b, rest = k.split(".", 3)
        mod = rest[: -len(".weight")]
        if mod in _LORA_CLIP_MAP:
            m[f"lora_te_text_model_encoder_layers_{b}_{_LORA_CLIP_MAP[mod]}"] = ("clip", k)
            m[f"lora_te1_text_model_encoder_layers_{b}_{_LORA_CLIP_MAP[mod]}"] = ("clip", k)
            m[f"text_encoder.text_model.encoder.layers.{b}.{mod}"] = ("clip", k)
    return m


def merge_lora(parts: Dict[str, Dict[str, torch.Tensor]], lora: Mapping[str, torch.Tensor], strength_model: float = 1.0,
               strength_clip: float = 1.0) -> int:
    """Fold a LoRA into `parts` in place; returns the number of weights patched.  Unknown LoRA modules are ignored, like
    the reference's load_lora (it only walks the keys it can map)."""
    keymap = lora_key_map(parts)
    n = 0
    for mod, (part, wkey) in keymap.items():
        up_k, down_k = mod + ".lora_up.weight", mod + ".lora_down.weight"
        if up_k not in lora or down_k not in lora:
            continue
        strength = strength_model if part == "unet" else strength_clip
        if strength == 0.0:
            continue
        up, down = lora[up_k].float(), lora[down_k].float()
        alpha = strength
        a_k = mod + ".alpha"
        if a_k in lora:
            alpha *= float(lora[a_k].item()) / down.shape[0]
        w = parts[part][wkey]
        delta = (alpha * torch.mm(up.flatten(start_dim=1), down.flatten(start_dim=1))).reshape(w.shape)
        # patch_weight_to_device (ModelPatcher.py:267-300): the sum is formed on an fp32 copy and rounded once
        parts[part][wkey] = (w.to(torch.float32) + delta).to(w.dtype)
        n += 1
    return n


def load_sd15(engine, path: str, lora_path: Optional[str] = None, strength_model: float = 1.0,
              strength_clip: float = 1.0) -> Dict[str, int]:
    """Checkpoint file (+ optional LoRA) -> engine.  Returns the tensor count loaded per part."""
    parts = split_sd15_checkpoint(load_state_dict_file(path))
    if lora_path:
        merge_lora(parts, load_state_dict_file(lora_path), strength_model, strength_clip)
    if parts["unet"]:
        engine.load_unet(parts["unet"])
    if parts["vae"]:
        engine.load_vae(parts["vae"])
    if parts["clip"]:
        engine.load_clip(parts["clip"])
    return {k: len(v) for k, v in parts.items()}


def load_flux_files(engine, unet_path: str, ae_path: Optional[str] = None, clip_l_path: Optional[str] = None,
                    t5_path: Optional[str] = None) -> Dict[str, int]:
    """The four files of the reference's Flux branch (src/user/pipeline.py:225-237: UnetLoaderGGUF "flux1-dev-Q8_0.gguf",
    VAELoader "ae.safetensors", DualCLIPLoaderGGUF "clip_l.safetensors" + "t5-v1_1-xxl-encoder-Q8_0.gguf") -> engine.
    GGUF or safetensors for the DiT and T5; Q8_0 is dequantised once at ingest.  Returns the tensor count loaded per part."""
    n: Dict[str, int] = {}
    dit = load_gguf(unet_path) if unet_path.endswith(".gguf") else _strip_optional(load_state_dict_file(unet_path), "model.diffusion_model.")
    engine.load_flux(dit)
    n["flux"] = len(dit)
    if ae_path:
        ae = {k: v for k, v in load_state_dict_file(ae_path).items() if k.startswith(("decoder.", "encoder."))}
        if not ae:
            raise ValueError(f"{ae_path}: no decoder.* / encoder.* tensors (not a Flux autoencoder file)")
        engine.load_vae(ae)
        n["vae"] = len(ae)
    if clip_l_path:
        sd = load_state_dict_file(clip_l_path)
        clip = _strip_optional(_strip_optional(sd, "cond_stage_model.transformer."), "text_model.")
        clip = {k: v for k, v in clip.items() if k.startswith(("embeddings.", "encoder.", "final_layer_norm."))}
        if "embeddings.token_embedding.weight" not in clip:
            raise ValueError(f"{clip_l_path}: no text_model.embeddings.token_embedding.weight (not a CLIP-L text encoder file)")
        clip.pop("embeddings.position_ids", None)
        engine.load_clip(clip)
        n["clip"] = len(clip)
    if t5_path:
        t5 = load_t5_gguf(t5_path) if t5_path.endswith(".gguf") else load_state_dict_file(t5_path)
        engine.load_t5(t5)
        n["t5"] = len(t5)
    return n


def _strip_optional(sd: Mapping[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    """Drop `prefix` from the keys that carry it (files come with or without the wrapping module's name)."""
    if not any(k.startswith(prefix) for k in sd):
        return dict(sd)
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
