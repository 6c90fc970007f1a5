"""Host-side boundary work of the Flux.1 path (reference: src/BlackForest/Flux.py): 2x2 patchify / unpatchify of the
16-channel latent (Flux3.forward :732-778), the rotary table of EmbedND (:36-64, 85-113) for the concatenated
(text, image) position ids, and the seeded synthetic weight layout used by benchmarks and tests."""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch

FLUX_DEV = dict(in_channels=16, vec_in_dim=768, context_in_dim=4096, hidden_size=3072, mlp_ratio=4.0, num_heads=24,
                depth=19, depth_single_blocks=38, axes_dim=(16, 56, 56), theta=10000, guidance_embed=True)


def patchify(x: torch.Tensor) -> Tuple[torch.Tensor, int, int]:
    """[B,c,h,w] -> tokens [B,(h/2)(w/2),4c] ("b c (h ph) (w pw) -> b (h w) (c ph pw)"); odd sizes are padded circularly to
    even ones like pad_to_patch_size (Flux.py:474-499)."""
    B, c, h, w = x.shape
    if h % 2 or w % 2:
        x = torch.nn.functional.pad(x, (0, w % 2, 0, h % 2), mode="circular")
        h, w = x.shape[2], x.shape[3]
    hl, wl = h // 2, w // 2
    img = x.reshape(B, c, hl, 2, wl, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, hl * wl, c * 4)
    return img.contiguous(), hl, wl


def unpatchify(out: torch.Tensor, hl: int, wl: int) -> torch.Tensor:
    B = out.shape[0]
    c = out.shape[-1] // 4
    return out.reshape(B, hl, wl, c, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(B, c, hl * 2, wl * 2)


def rope_table(hl: int, wl: int, n_txt: int, axes_dim: Sequence[int] = (16, 56, 56), theta: int = 10000) -> torch.Tensor:
    """(cos, sin) per token and rotation pair, [n_txt + hl*wl, sum(axes)/2, 2] fp32: text ids are all zero, image ids are
    (0, row, col); omega is evaluated in float64 as the reference does."""
    ids = torch.zeros(n_txt + hl * wl, 3)
    grid = torch.zeros(hl, wl, 3)
    grid[..., 1] = torch.arange(hl, dtype=torch.float32)[:, None]
    grid[..., 2] = torch.arange(wl, dtype=torch.float32)[None, :]
    ids[n_txt:] = grid.reshape(-1, 3)
    outs = []
    for i, d in enumerate(axes_dim):
        scale = torch.linspace(0, (d - 2) / d, steps=d // 2, dtype=torch.float64)
        omega = 1.0 / (theta ** scale)
        ang = torch.einsum("n,d->nd", ids[:, i].to(torch.float32), omega)
        outs.append(torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1))
    return torch.cat(outs, dim=-2).to(torch.float32)


def flux_shapes(cfg=FLUX_DEV) -> Dict[str, Tuple[int, ...]]:
    """State-dict layout of Flux3 (Flux.py:548-656)."""
    C, H = cfg["hidden_size"], cfg["num_heads"]
    hd, M = C // H, int(C * cfg["mlp_ratio"])
    s: Dict[str, Tuple[int, ...]] = {}

    def lin(p, o, i):
        s[p + ".weight"] = (o, i)
        s[p + ".bias"] = (o,)

    lin("img_in", C, cfg["in_channels"] * 4)
    for n, i in (("time_in", 256), ("vector_in", cfg["vec_in_dim"]), ("guidance_in", 256)):
        if n == "guidance_in" and not cfg["guidance_embed"]:
            continue
        lin(n + ".in_layer", C, i)
        lin(n + ".out_layer", C, C)
    lin("txt_in", C, cfg["context_in_dim"])
    for b in range(cfg["depth"]):
        p = f"double_blocks.{b}"
        for t in ("img", "txt"):
            lin(f"{p}.{t}_mod.lin", 6 * C, C)
            lin(f"{p}.{t}_attn.qkv", 3 * C, C)
            s[f"{p}.{t}_attn.norm.query_norm.scale"] = (hd,)
            s[f"{p}.{t}_attn.norm.key_norm.scale"] = (hd,)
            lin(f"{p}.{t}_attn.proj", C, C)
            lin(f"{p}.{t}_mlp.0", M, C)
            lin(f"{p}.{t}_mlp.2", C, M)
    for b in range(cfg["depth_single_blocks"]):
        p = f"single_blocks.{b}"
        lin(f"{p}.linear1", 3 * C + M, C)
        lin(f"{p}.linear2", C, C + M)
        s[f"{p}.norm.query_norm.scale"] = (hd,)
        s[f"{p}.norm.key_norm.scale"] = (hd,)
        lin(f"{p}.modulation.lin", 3 * C, C)
    lin("final_layer.linear", cfg["in_channels"] * 4, C)
    lin("final_layer.adaLN_modulation.1", 2 * C, C)
    return s


def infer_config(sd) -> Dict[str, object]:
    """Flux3 hyper-parameters read off a state dict's shapes (what detect_unet_config does for the reference's loader):
    everything but the rotary axes / theta, which are not stored in the weights."""
    def need(k):
        if k not in sd:
            raise ValueError(f"Flux state dict: {k} is missing")
        return tuple(int(v) for v in sd[k].shape)

    C, cin4 = need("img_in.weight")
    hd = need("double_blocks.0.img_attn.norm.query_norm.scale")[0]
    M = need("double_blocks.0.img_mlp.0.weight")[0]
    depth = 0
    while f"double_blocks.{depth}.img_attn.qkv.weight" in sd:
        depth += 1
    single = 0
    while f"single_blocks.{single}.linear1.weight" in sd:
        single += 1
    if C % hd != 0 or cin4 % 4 != 0:
        raise ValueError(f"Flux state dict: hidden size {C} / head width {hd} / patch width {cin4} are inconsistent")
    return dict(in_channels=cin4 // 4, vec_in_dim=need("vector_in.in_layer.weight")[1], context_in_dim=need("txt_in.weight")[1],
                hidden_size=C, mlp_ratio=M / C, num_heads=C // hd, depth=depth, depth_single_blocks=single,
                guidance_embed="guidance_in.in_layer.weight" in sd)


def validate_state_dict(sd) -> Dict[str, object]:
    """Names and shapes of a Flux3 state dict against the layout the engine consumes; returns the inferred configuration.
    Raises ValueError naming what is wrong (no fallback for a malformed checkpoint).  The joint attention kernel is built
    for 128-wide heads."""
    cfg = infer_config(sd)
    want = flux_shapes(cfg)
    missing = sorted(k for k in want if k not in sd)
    wrong = sorted(k for k in want if k in sd and tuple(sd[k].shape) != want[k])
    if missing or wrong:
        raise ValueError(f"Flux state dict does not match the Flux3 layout: missing {missing[:4]}{'...' if len(missing) > 4 else ''}, "
                         f"wrong shape {[(k, tuple(sd[k].shape), want[k]) for k in wrong[:3]]}")
    if cfg["hidden_size"] // cfg["num_heads"] != 128:
        raise ValueError(f"Flux: head width {cfg['hidden_size'] // cfg['num_heads']} is not supported (128 expected)")
    return cfg
