"""Batch-of-prompts data parallelism over the GPUs of one box (one process per GPU, torch.distributed).

The reference has no multi-GPU path; the sampler trajectory of an image depends only on its own noise
and conditioning, so images shard across ranks with NO per-step collective.  Collectives used, all outside the loop:
  * broadcast of the weights from rank 0 (once),
  * none for the noise: every rank replays the reference's full-batch draw (one seeded CPU generator over [B,4,h,w],
    src/sample/ksampler_util.py:287-295) and keeps its slice, so a sharded run reproduces the single-GPU batch result
    (scatter_rows / gather_rows remain for data that only rank 0 holds, e.g. per-image prompts or input images),
  * gather of the final latents (256 KB per 1024^2 image) or decoded images.
Backend: "nccl" on GPUs (NVLink/NVSwitch), "gloo" in the CPU tests.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Rank r owns images [lo, hi): consecutive, sizes differ by at most one (ragged batches allowed)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_state_dict(shapes: Dict[str, Tuple[int, ...]], make_tensor, device, dtype=torch.float16,
                         src: int = 0) -> Dict[str, torch.Tensor]:
    """Rank `src` materialises each tensor with make_tensor(name, shape); everyone receives it."""
    out = {}
    rank = dist.get_rank() if dist.is_initialized() else 0
    for name in sorted(shapes):
        if rank == src:
            t = make_tensor(name, shapes[name]).to(device=device, dtype=dtype).contiguous()
        else:
            t = torch.empty(shapes[name], dtype=dtype, device=device)
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.broadcast(t, src)
        out[name] = t
    return out


def scatter_rows(full: Optional[torch.Tensor], shape_tail: Tuple[int, ...], batch: int, device, dtype=torch.float32,
                 src: int = 0) -> torch.Tensor:
    """Scatter rows of `full` ([batch, *shape_tail], present on `src`) according to shard_range."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return full.to(device)
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(batch, rank, world)
    mine = torch.empty((hi - lo,) + tuple(shape_tail), dtype=dtype, device=device)
    # ragged-safe: point-to-point from src (sizes may differ per rank)
    if rank == src:
        reqs = []
        for r in range(world):
            a, b = shard_range(batch, r, world)
            part = full[a:b].to(device=device, dtype=dtype).contiguous()
            if r == src:
                mine.copy_(part)
            elif b > a:
                reqs.append(dist.isend(part, r))
        for q in reqs:
            q.wait()
    elif hi > lo:
        dist.recv(mine, src)
    return mine


def gather_rows(mine: torch.Tensor, batch: int, dst: int = 0) -> Optional[torch.Tensor]:
    """Inverse of scatter_rows: returns the [batch, ...] tensor on `dst`, None elsewhere."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return mine
    rank, world = dist.get_rank(), dist.get_world_size()
    if rank == dst:
        parts: List[torch.Tensor] = []
        for r in range(world):
            a, b = shard_range(batch, r, world)
            if r == dst:
                parts.append(mine)
            else:
                buf = torch.empty((b - a,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
                if b > a:
                    dist.recv(buf, r)
                parts.append(buf)
        return torch.cat(parts)
    if mine.shape[0] > 0:
        dist.send(mine.contiguous(), dst)
    return None


def sample_sharded(engine, seed: int, steps: int, cfg: float, sampler_name: str, scheduler: str, positive: torch.Tensor,
                   negative: torch.Tensor, latent_image: Dict[str, torch.Tensor], images_per_call: int = 0, **kw):
    """KSampler.sample for a batch sharded over the ranks.  positive/negative: [1 or B, T, 768] on every rank.
    Returns ({"samples": [B,4,h,w]},) on rank 0 and (None,) elsewhere.  Reproduces the single-process batch result for
    every sampler: the initial noise is each rank's slice of the same full-batch draw, and the per-step noise of the ancestral
    / SDE samplers is drawn on every rank for the WHOLE batch from identically seeded generators and sliced (`batch_slice`),
    so image i gets the noise it would get as row i of the unsharded batch -- never the same noise as another image.
    images_per_call > 0 bounds the UNet batch: a rank's images are sampled in consecutive groups of that size (one launch
    program, bounded activation memory); for the deterministic samplers the result is unchanged, for the ancestral / SDE
    ones every group re-seeds the generators (each group consumes the whole-batch noise stream from its start)."""
    from . import sampling as S

    latent = latent_image["samples"]
    B = latent.shape[0]
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(B, rank, world)
    # Every rank replays the reference's prepare_noise (ksampler_util.py:274-295) for the FULL batch: one CPU generator
    # seeded with `seed` (torch.manual_seed also seeds the CUDA generators).  It costs microseconds, needs no scatter, and
    # leaves every generator of every rank in exactly the state the single-process batch run has at this point -- which is
    # what makes the per-step noise below shard-invariant.
    noise = S.prepare_noise(latent, seed)[lo:hi]
    dev = engine.device
    out = None
    if hi > lo:
        step = images_per_call if images_per_call > 0 else hi - lo
        parts = []
        for a in range(lo, hi, step):
            b = min(hi, a + step)
            if a > lo:
                S.prepare_noise(latent, seed)  # same generator state at the start of every group
            pos = positive if positive.shape[0] == 1 else positive[a:b]
            neg = negative if negative.shape[0] == 1 else negative[a:b]
            res = S.sample(engine, seed, steps, cfg, sampler_name, scheduler, pos, neg, {"samples": latent[a:b]},
                           noise=noise[a - lo:b - lo], batch_slice=(a, b, B), **kw)
            parts.append(res[0]["samples"].to(dev))
        out = torch.cat(parts) if len(parts) > 1 else parts[0]
    else:
        out = torch.empty((0,) + tuple(latent.shape[1:]), device=dev)
    full = gather_rows(out, B)
    return ({"samples": full.cpu()} if full is not None else None,)


def sample_flux_sharded(engine, seed: int, steps: int, positive, negative, latent_image: Dict[str, torch.Tensor], **kw):
    """flux_sampling.sample_flux for a batch of latents sharded over the ranks (same scheme as sample_sharded: rank 0 draws
    the full-batch noise as the reference does, slices travel point to point, no collective inside the loop, latents are
    gathered on rank 0).  positive / negative: (T5 states [1,Nt,C], pooled vector [1,V]) on every rank.  `euler_cfgpp`
    is deterministic given the noise, so the sharded batch reproduces the single-process batch."""
    from . import flux_sampling as FS
    from . import sampling as S

    latent = latent_image["samples"]
    B = latent.shape[0]
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(B, rank, world)
    noise_full = S.prepare_noise(latent, seed) if rank == 0 else None
    dev = engine.device
    noise = scatter_rows(noise_full, tuple(latent.shape[1:]), B, dev)
    if hi > lo:
        res = FS.sample_flux(engine, seed, steps, positive, negative, {"samples": latent[lo:hi]}, noise=noise.cpu(), **kw)
        out = res[0]["samples"].to(dev)
    else:
        out = torch.empty((0,) + tuple(latent.shape[1:]), device=dev)
    full = gather_rows(out, B)
    return ({"samples": full.cpu()} if full is not None else None,)
