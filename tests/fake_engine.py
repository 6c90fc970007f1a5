"""CPU stand-in for lightdiffusion_next_b200.engine.Engine used only to test HOST logic (sampler loop, sharding, seams,
pipeline composition) without a GPU: every device call is answered by the oracle / plain torch.  Test infrastructure --
never a product path."""
import torch

from lightdiffusion_next_b200.schedule import DiscreteSchedule
from oracle import sd15_oracle as O


class FakeEngine:
    def __init__(self, sd, vae_sd=None, clip_sd=None, unet_cfg=None):
        self.sd = sd
        self.unet_cfg = unet_cfg or O.SD15  # host-logic tests that pin nothing to a golden use a narrow UNet (conftest.TINY_UNET)
        self.vae_sd = vae_sd
        self.clip_sd = clip_sd
        self.device = torch.device("cpu")
        self.schedule = DiscreteSchedule()
        self.tables = O.make_sigma_tables()
        self.ctx = None
        self.denoise_calls = 0
        self.context_uploads = 0
        self.weights_epoch = {}
        self.extra = []

    def set_context(self, ctx):
        self.ctx = ctx.clone()
        self.context_uploads += 1

    def denoise(self, x, sigma, out=None):
        r = O.apply_model(self.sd, x, sigma, self.ctx, self.tables, self.unet_cfg)
        self.denoise_calls += 1
        if out is not None:
            out.copy_(r)
            return out
        return r

    def resample_bilinear(self, x, size, out=None):
        r = torch.nn.functional.interpolate(x, size=tuple(size), mode="bilinear", align_corners=False)
        if out is not None:
            out.copy_(r)
            return out
        return r

    def cfg_step(self, x, du, dc, cfg, mode, c0=0.0, c1=0.0, c2=0.0, noise=None, x_out=None, denoised_out=None):
        den = torch.lerp(du, dc, cfg)
        if denoised_out is not None:
            denoised_out.copy_(den)
        if mode == 0:
            x_out.copy_(c0 * x - c1 * den)
        elif mode == 1:
            r = x + ((x - den) / c2) * c0
            if noise is not None:
                r = r + noise * c1
            x_out.copy_(r)

    # ---- CLIP / VAE (pipeline composition tests)
    def clip_vocab(self):
        return int(self.clip_sd["embeddings.token_embedding.weight"].shape[0])

    def set_clip_extra_embeddings(self, vectors):
        self.extra = list(vectors)

    def clip_encode(self, ids):
        sd = self.clip_sd
        if self.extra:
            tok = sd["embeddings.token_embedding.weight"]
            sd = dict(sd)
            sd["embeddings.token_embedding.weight"] = torch.cat([tok, torch.stack(self.extra).to(tok.dtype)])
        return O.clip_encode(sd, ids)

    def vae_decode(self, z):
        return O.vae_decode(self.vae_sd, z)
