"""CPU stand-in for lightdiffusion_next_b200.engine.Engine used only to test HOST logic (sampler loop, sharding)
without a GPU: denoise() is answered by the oracle, cfg_step() by plain torch."""
import torch

from lightdiffusion_next_b200.schedule import DiscreteSchedule
from oracle import sd15_oracle as O


class FakeEngine:
    def __init__(self, sd):
        self.sd = sd
        self.device = torch.device("cpu")
        self.schedule = DiscreteSchedule()
        self.tables = O.make_sigma_tables()
        self.ctx = None
        self.denoise_calls = 0

    def set_context(self, ctx):
        self.ctx = ctx.clone()

    def denoise(self, x, sigma, out=None):
        r = O.apply_model(self.sd, x, sigma, self.ctx, self.tables)
        self.denoise_calls += 1
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
