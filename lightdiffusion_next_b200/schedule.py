"""Host-side sigma schedules and solver coefficients (fp32 torch on CPU, same op order as the reference so the
trajectories line up).  Mirrors: ModelSamplingDiscrete (src/sample/sampling.py:221-356), calculate_sigmas
(src/sample/ksampler_util.py:244-271), get_sigmas_karras / get_ancestral_step (src/sample/sampling_util.py:106-151)."""
from __future__ import annotations

import math

import torch


class DiscreteSchedule:
    """SD1.x scaled-linear beta schedule, 1000 steps (linear_start 0.00085, linear_end 0.012)."""

    def __init__(self, linear_start: float = 0.00085, linear_end: float = 0.012, timesteps: int = 1000):
        betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2
        alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        sigmas = ((1 - alphas_cumprod) / alphas_cumprod) ** 0.5
        self.sigmas = sigmas.float()
        self.log_sigmas = sigmas.log().float()

    @property
    def sigma_min(self) -> torch.Tensor:
        return self.sigmas[0]

    @property
    def sigma_max(self) -> torch.Tensor:
        return self.sigmas[-1]

    def timestep(self, sigma: torch.Tensor) -> torch.Tensor:
        dists = sigma.log() - self.log_sigmas[:, None]
        return dists.abs().argmin(dim=0).view(sigma.shape)

    def sigma(self, timestep: torch.Tensor) -> torch.Tensor:
        t = torch.clamp(timestep.float(), min=0, max=len(self.sigmas) - 1)
        low, high, w = t.floor().long(), t.ceil().long(), t.frac()
        return ((1 - w) * self.log_sigmas[low] + w * self.log_sigmas[high]).exp()


class FluxSchedule:
    """ModelSamplingFlux (src/sample/sampling.py:172-218): flow-matching time shift sigma(t) = e^mu / (e^mu + (1/t - 1)),
    mu = shift = 1.15 by default, tabulated at t = 1/10000 ... 1; the model's timestep IS sigma. Duck-types
    DiscreteSchedule for calculate_sigmas (the reference's Flux pipeline uses the `beta` scheduler)."""

    def __init__(self, shift: float = 1.15, timesteps: int = 10000):
        self.shift = shift
        self.sigmas = self.sigma(torch.arange(1, timesteps + 1, 1) / timesteps)

    @property
    def sigma_min(self) -> torch.Tensor:
        return self.sigmas[0]

    @property
    def sigma_max(self) -> torch.Tensor:
        return self.sigmas[-1]

    def timestep(self, sigma: torch.Tensor) -> torch.Tensor:
        return sigma

    def sigma(self, timestep: torch.Tensor) -> torch.Tensor:
        return math.exp(self.shift) / (math.exp(self.shift) + (1 / timestep - 1) ** 1.0)


def get_sigmas_karras(n: int, sigma_min: float, sigma_max: float, rho: float = 7.0) -> torch.Tensor:
    ramp = torch.linspace(0, 1, n)
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return torch.cat([sigmas, sigmas.new_zeros([1])])


def normal_scheduler(ms: DiscreteSchedule, steps: int) -> torch.Tensor:
    start = ms.timestep(ms.sigma_max)
    end = ms.timestep(ms.sigma_min)
    timesteps = torch.linspace(start, end, steps)
    sigs = [float(ms.sigma(timesteps[i])) for i in range(len(timesteps))]
    return torch.FloatTensor(sigs + [0.0])


def simple_scheduler(ms: DiscreteSchedule, steps: int) -> torch.Tensor:
    """Every (1000 / steps)-th discrete sigma, walking down from sigma_max (ksampler_util.py:180-199)."""
    n = len(ms.sigmas)
    ss = n / steps
    sigs = [float(ms.sigmas[-(1 + int(x * ss))]) for x in range(steps)]
    return torch.FloatTensor(sigs + [0.0])


def beta_scheduler(ms: DiscreteSchedule, steps: int, alpha: float = 0.6, beta: float = 0.6) -> torch.Tensor:
    """Timesteps at the Beta(alpha, beta) quantiles of an even grid, duplicates dropped (arXiv 2407.12173;
    ksampler_util.py:202-241): the schedule may hold fewer than `steps` sigmas."""
    import numpy as np
    import scipy.stats

    total = len(ms.sigmas) - 1
    ts = scipy.stats.beta.ppf(1 - np.linspace(0, 1, steps, endpoint=False), alpha, beta)
    idx = np.rint(ts * total).astype(np.int32)
    uniq, first = np.unique(idx, return_index=True)
    ordered = uniq[np.argsort(first)]
    return torch.FloatTensor([float(ms.sigmas[i]) for i in ordered] + [0.0])


SCHEDULERS = ("karras", "normal", "simple", "beta")


def calculate_sigmas(ms: DiscreteSchedule, scheduler_name: str, steps: int) -> torch.Tensor:
    if scheduler_name == "karras":
        return get_sigmas_karras(steps, float(ms.sigma_min), float(ms.sigma_max))
    if scheduler_name == "normal":
        return normal_scheduler(ms, steps)
    if scheduler_name == "simple":
        return simple_scheduler(ms, steps)
    if scheduler_name == "beta":
        return beta_scheduler(ms, steps)
    raise ValueError(f"unsupported scheduler {scheduler_name!r} ({' | '.join(SCHEDULERS)})")


def get_ancestral_step(sigma_from: torch.Tensor, sigma_to: torch.Tensor, eta: float = 1.0):
    sigma_up = min(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
    return sigma_down, sigma_up


def max_denoise(ms: DiscreteSchedule, sigmas: torch.Tensor) -> bool:
    max_sigma = float(ms.sigma_max)
    sigma = float(sigmas[0])
    return math.isclose(max_sigma, sigma, rel_tol=1e-05) or sigma > max_sigma
