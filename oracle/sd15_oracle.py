"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path (lightdiffusion_next_b200/*).

Plain-PyTorch fp32 CPU restatement of the LightDiffusion-Next SD1.5 sampling hot path, written from the
reference's behaviour (not its code) so that the CUDA engine can be checked against it on machines where
/root/reference does not exist.  Each function cites the reference file:line it follows (paths relative to the
reference repo).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.

Parity pinning: the reference ships no tests and no golden vectors.  This oracle is pinned
against outputs of the reference itself, generated in the build container by tests/golden/make_golden.py
(which imports /root/reference read-only) and committed under tests/golden/*.pt; tests/test_oracle_golden.py
checks every function here against those fixtures.
"""
from __future__ import annotations

import math
import zlib
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# ----------------------------------------------------------------------------------------------------------------
# SD1.5 architecture constants (src/SD15/SD15.py:17-28, src/NeuralNetwork/unet.py:941-1080)
# ----------------------------------------------------------------------------------------------------------------
SD15 = dict(
    in_channels=4, out_channels=4, model_channels=320, channel_mult=(1, 2, 4, 4), num_res_blocks=2,
    attn_levels=(True, True, True, False), num_heads=8, context_dim=768, time_embed_dim=1280,
)
LATENT_SCALE = 0.18215  # src/Utilities/Latent.py:41-62


# ----------------------------------------------------------------------------------------------------------------
# Discrete sigma schedule  (ModelSamplingDiscrete, src/sample/sampling.py:221-356; make_beta_schedule,
# src/sample/sampling_util.py:18-39)
# ----------------------------------------------------------------------------------------------------------------
def make_sigma_tables() -> Tuple[Tensor, Tensor]:
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2
    alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
    sigmas = ((1 - alphas_cumprod) / alphas_cumprod) ** 0.5
    return sigmas.float(), sigmas.log().float()


def timestep_index(sigma: Tensor, log_sigmas: Tensor) -> Tensor:
    """sampling.py:309-320 — nearest discrete index of sigma in log space (no interpolation)."""
    dists = sigma.log() - log_sigmas[:, None]
    return dists.abs().argmin(dim=0).view(sigma.shape)


def sigma_of_timestep(t: Tensor, log_sigmas: Tensor) -> Tensor:
    """sampling.py:322-338 — log-linear interpolation of the table."""
    t = torch.clamp(t.float(), min=0, max=len(log_sigmas) - 1)
    low, high, w = t.floor().long(), t.ceil().long(), t.frac()
    return ((1 - w) * log_sigmas[low] + w * log_sigmas[high]).exp()


def sigmas_karras(n: int, sigma_min: float, sigma_max: float, rho: float = 7.0) -> Tensor:
    """sampling_util.py:106-125."""
    ramp = torch.linspace(0, 1, n)
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    sig = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return torch.cat([sig, sig.new_zeros([1])])


def sigmas_normal(steps: int, sigmas: Tensor, log_sigmas: Tensor) -> Tensor:
    """ksampler_util.py:152-177 (sgm=False)."""
    start = timestep_index(sigmas[-1], log_sigmas)
    end = timestep_index(sigmas[0], log_sigmas)
    ts = torch.linspace(start, end, steps)
    out = [float(sigma_of_timestep(ts[i], log_sigmas)) for i in range(len(ts))]
    return torch.FloatTensor(out + [0.0])


def calculate_sigmas(scheduler: str, steps: int) -> Tensor:
    """ksampler_util.py:244-271."""
    sigmas, log_sigmas = make_sigma_tables()
    if scheduler == "karras":
        return sigmas_karras(steps, float(sigmas[0]), float(sigmas[-1]))
    if scheduler == "normal":
        return sigmas_normal(steps, sigmas, log_sigmas)
    if scheduler == "simple":  # simple_scheduler, ksampler_util.py:180-199
        ss = len(sigmas) / steps
        return torch.FloatTensor([float(sigmas[-(1 + int(x * ss))]) for x in range(steps)] + [0.0])
    if scheduler == "beta":  # beta_scheduler (alpha = beta = 0.6), ksampler_util.py:202-241
        import numpy as np
        import scipy.stats
        ts = scipy.stats.beta.ppf(1 - np.linspace(0, 1, steps, endpoint=False), 0.6, 0.6)
        idx = np.rint(ts * (len(sigmas) - 1)).astype(np.int32)
        uniq, first = np.unique(idx, return_index=True)
        return torch.FloatTensor([float(sigmas[i]) for i in uniq[np.argsort(first)]] + [0.0])
    raise ValueError(scheduler)


def timestep_embedding(t: Tensor, dim: int, max_period: int = 10000) -> Tensor:
    """sampling_util.py:56-76."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


# ----------------------------------------------------------------------------------------------------------------
# UNet block table + parameter shapes (LDM constructor rules, unet.py:344-677) and seeded synthetic weights
# ----------------------------------------------------------------------------------------------------------------
def unet_blocks(cfg=SD15):
    """Returns (input_blocks, middle, output_blocks): lists of blocks; a block is a list of
    ('conv_in'|'res'|'st'|'down'|'up', prefix, cin, cout)."""
    mc, mult, nres = cfg["model_channels"], cfg["channel_mult"], cfg["num_res_blocks"]
    attn = cfg["attn_levels"]
    inp = [[("conv_in", "input_blocks.0.0", cfg["in_channels"], mc)]]
    chans = [mc]
    ch, bi = mc, 1
    for level, m in enumerate(mult):
        for _ in range(nres):
            blk = [("res", f"input_blocks.{bi}.0", ch, m * mc)]
            ch = m * mc
            if attn[level]:
                blk.append(("st", f"input_blocks.{bi}.1", ch, ch))
            inp.append(blk)
            chans.append(ch)
            bi += 1
        if level != len(mult) - 1:
            inp.append([("down", f"input_blocks.{bi}.0.op", ch, ch)])
            chans.append(ch)
            bi += 1
    mid = [("res", "middle_block.0", ch, ch), ("st", "middle_block.1", ch, ch), ("res", "middle_block.2", ch, ch)]
    out = []
    oi = 0
    for level in reversed(range(len(mult))):
        for i in range(nres + 1):
            ich = chans.pop()
            blk = [("res", f"output_blocks.{oi}.0", ch + ich, mc * mult[level])]
            ch = mc * mult[level]
            sub = 1
            if attn[level]:
                blk.append(("st", f"output_blocks.{oi}.{sub}", ch, ch))
                sub += 1
            if level > 0 and i == nres:
                blk.append(("up", f"output_blocks.{oi}.{sub}.conv", ch, ch))
            out.append(blk)
            oi += 1
    return inp, mid, out


def unet_param_shapes(cfg=SD15) -> Dict[str, Tuple[int, ...]]:
    """State-dict keys (without the 'model.diffusion_model.' prefix) -> shapes."""
    mc, te, cd = cfg["model_channels"], cfg["time_embed_dim"], cfg["context_dim"]
    s: Dict[str, Tuple[int, ...]] = {}

    def lin(p, o, i, bias=True):
        s[p + ".weight"] = (o, i)
        if bias:
            s[p + ".bias"] = (o,)

    def conv(p, o, i, k):
        s[p + ".weight"] = (o, i, k, k)
        s[p + ".bias"] = (o,)

    def norm(p, c):
        s[p + ".weight"] = (c,)
        s[p + ".bias"] = (c,)

    lin("time_embed.0", te, mc)
    lin("time_embed.2", te, te)
    inp, mid, out = unet_blocks(cfg)
    for blk in inp + [mid] + out:
        for kind, p, cin, cout in blk:
            if kind == "conv_in":
                conv(p, cout, cin, 3)
            elif kind == "res":
                norm(p + ".in_layers.0", cin)
                conv(p + ".in_layers.2", cout, cin, 3)
                lin(p + ".emb_layers.1", cout, te)
                norm(p + ".out_layers.0", cout)
                conv(p + ".out_layers.3", cout, cout, 3)
                if cin != cout:
                    conv(p + ".skip_connection", cout, cin, 1)
            elif kind == "st":
                c = cin
                norm(p + ".norm", c)
                conv(p + ".proj_in", c, c, 1)
                tb = p + ".transformer_blocks.0"
                for a, kd in (("attn1", c), ("attn2", cd)):
                    lin(f"{tb}.{a}.to_q", c, c, bias=False)
                    lin(f"{tb}.{a}.to_k", c, kd, bias=False)
                    lin(f"{tb}.{a}.to_v", c, kd, bias=False)
                    lin(f"{tb}.{a}.to_out.0", c, c)
                lin(f"{tb}.ff.net.0.proj", 8 * c, c)
                lin(f"{tb}.ff.net.2", c, 4 * c)
                for n in ("norm1", "norm2", "norm3"):
                    norm(f"{tb}.{n}", c)
                conv(p + ".proj_out", c, c, 1)
            elif kind in ("down", "up"):
                conv(p, cout, cin, 3)
    norm("out.0", mc)
    conv("out.2", cfg["out_channels"], mc, 3)
    return s


_RESIDUAL_OUT = ("out_layers.3.", "proj_out.", "to_out.0.", "ff.net.2.", "conv2.", "out_proj.", "mlp.fc2.")


def synth_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 1234, dtype=torch.float16) -> Dict[str, Tensor]:
    """Seeded synthetic weights, independent of iteration order (per-tensor seed = crc32(name) ^ seed).
    Matrices ~ N(0, 1/fan_in) (x0.5 on residual-branch output layers), norm gains ~ 1 + 0.1 N(0,1), biases and
    norm shifts ~ 0.05 N(0,1); rounded to `dtype` (the reference stores SD1.5 weights as fp16, unet.py:1127-1146)."""
    sd = {}
    for name, shape in shapes.items():
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ seed) & 0x7FFFFFFF)
        if len(shape) > 1:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            if "embedding" in name:
                w = torch.randn(shape, generator=g) * 0.02
            else:
                w = torch.randn(shape, generator=g) / math.sqrt(fan_in)
                if any(k in name for k in _RESIDUAL_OUT):
                    w = w * 0.5
        elif name.endswith(".weight"):
            w = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            w = 0.05 * torch.randn(shape, generator=g)
        sd[name] = w.to(dtype)
    return sd


# ----------------------------------------------------------------------------------------------------------------
# UNet forward (UNetModel1.forward, unet.py:679-770)
# ----------------------------------------------------------------------------------------------------------------
def _w(sd, k):
    return sd[k].float()


def _gn(sd, p, x, eps):
    return F.group_norm(x, 32, _w(sd, p + ".weight"), _w(sd, p + ".bias"), eps)


def _conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, _w(sd, p + ".weight"), _w(sd, p + ".bias"), stride=stride, padding=padding)


def _linear(sd, p, x, bias=True):
    return F.linear(x, _w(sd, p + ".weight"), _w(sd, p + ".bias") if bias else None)


def resblock(sd, p, x, emb):
    """ResBlock1._forward, src/AutoEncoders/ResBlock.py:315-335 (GroupNorm eps 1e-5)."""
    h = _conv(sd, p + ".in_layers.2", F.silu(_gn(sd, p + ".in_layers.0", x, 1e-5)))
    h = h + _linear(sd, p + ".emb_layers.1", F.silu(emb))[:, :, None, None]
    h = _conv(sd, p + ".out_layers.3", F.silu(_gn(sd, p + ".out_layers.0", h, 1e-5)))
    if (p + ".skip_connection.weight") in sd:
        x = _conv(sd, p + ".skip_connection", x, padding=0)
    return x + h


def attention(q: Tensor, k: Tensor, v: Tensor, heads: int, mask: Optional[Tensor] = None) -> Tensor:
    """optimized_attention, src/Attention/AttentionMethods.py:107-134: softmax(q k^T / sqrt(d)) v per head."""
    b, n, c = q.shape
    d = c // heads
    q, k, v = (t.view(b, -1, heads, d).transpose(1, 2) for t in (q, k, v))
    if mask is None and n * k.shape[2] > (1 << 24):
        # same maths without materialising the N x N matrix (what the reference calls: F.scaled_dot_product_attention,
        # AttentionMethods.py:130) — only used at benchmark sizes where the explicit form would need > 10 GB
        return F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, n, c)
    s = torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5)
    if mask is not None:
        s = s + mask
    o = torch.matmul(torch.softmax(s, dim=-1), v)
    return o.transpose(1, 2).reshape(b, n, c)


def cross_attention(sd, p, x, context, heads):
    """CrossAttention.forward, src/Attention/Attention.py:100-124."""
    ctx = x if context is None else context
    q = _linear(sd, p + ".to_q", x, bias=False)
    k = _linear(sd, p + ".to_k", ctx, bias=False)
    v = _linear(sd, p + ".to_v", ctx, bias=False)
    return _linear(sd, p + ".to_out.0", attention(q, k, v, heads))


def _ln(sd, p, x):
    return F.layer_norm(x, x.shape[-1:], _w(sd, p + ".weight"), _w(sd, p + ".bias"), 1e-5)


def transformer_block(sd, p, x, context, heads):
    """BasicTransformerBlock._forward, src/NeuralNetwork/transformer.py:186-245; GEGLU Activation.py:30-31."""
    x = x + cross_attention(sd, p + ".attn1", _ln(sd, p + ".norm1", x), None, heads)
    x = x + cross_attention(sd, p + ".attn2", _ln(sd, p + ".norm2", x), context, heads)
    h = _linear(sd, p + ".ff.net.0.proj", _ln(sd, p + ".norm3", x))
    a, gate = h.chunk(2, dim=-1)
    return _linear(sd, p + ".ff.net.2", a * F.gelu(gate)) + x


def spatial_transformer(sd, p, x, context, heads):
    """SpatialTransformer.forward, transformer.py:342-377 (GroupNorm eps 1e-6, 1x1-conv projections)."""
    b, c, h, w = x.shape
    x_in = x
    x = _conv(sd, p + ".proj_in", _gn(sd, p + ".norm", x, 1e-6), padding=0)
    x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
    x = transformer_block(sd, p + ".transformer_blocks.0", x, context, heads)
    x = x.reshape(b, h, w, c).permute(0, 3, 1, 2)
    return _conv(sd, p + ".proj_out", x, padding=0) + x_in


def unet_forward(sd: Dict[str, Tensor], x: Tensor, t: Tensor, context: Tensor, cfg=SD15) -> Tensor:
    """x [B,4,h,w] fp32 (already scaled), t [B] integer-valued float, context [B,77k,768] -> eps [B,4,h,w]."""
    heads = cfg["num_heads"]
    emb = _linear(sd, "time_embed.2", F.silu(_linear(sd, "time_embed.0", timestep_embedding(t, cfg["model_channels"]))))
    inp, mid, out = unet_blocks(cfg)

    def run(blk, h):
        for kind, p, cin, cout in blk:
            if kind == "conv_in":
                h = _conv(sd, p, h)
            elif kind == "res":
                h = resblock(sd, p, h, emb)
            elif kind == "st":
                h = spatial_transformer(sd, p, h, context, heads)
            elif kind == "down":
                h = _conv(sd, p, h, stride=2)  # Downsample1, ResBlock.py:173-182
            elif kind == "up":
                h = _conv(sd, p, F.interpolate(h, scale_factor=2, mode="nearest"))  # Upsample1, ResBlock.py:106-138
        return h

    hs = []
    h = x.float()
    for blk in inp:
        h = run(blk, h)
        hs.append(h)
    h = run(mid, h)
    for blk in out:
        h = torch.cat([h, hs.pop()], dim=1)
        h = run(blk, h)
    return _conv(sd, "out.2", F.silu(_gn(sd, "out.0", h, 1e-5)))


def apply_model(sd, x: Tensor, sigma: Tensor, context: Tensor, tables=None, cfg=SD15) -> Tensor:
    """BaseModel.apply_model, src/Model/ModelBase.py:72-133 with EPS scaling (sampling.py:26-56):
    denoised = x - unet(x / sqrt(sigma^2+1), timestep(sigma), ctx) * sigma."""
    _, log_sigmas = tables or make_sigma_tables()
    s = sigma.view(-1, 1, 1, 1)
    xc = x / (s ** 2 + 1.0) ** 0.5
    t = timestep_index(sigma, log_sigmas).float()
    eps = unet_forward(sd, xc, t, context, cfg)
    return x - eps * s


# ----------------------------------------------------------------------------------------------------------------
# CFG + samplers
# ----------------------------------------------------------------------------------------------------------------
def cfg_denoise(model: Callable[[Tensor, Tensor, Tensor], Tensor], x: Tensor, sigma: Tensor, cond: Tensor,
                uncond: Tensor, cfg: float) -> Tensor:
    """calc_cond_batch + cfg_function (src/cond/cond.py:150-288, src/sample/CFG.py:6-83): one batched call with rows
    [uncond..., cond...] and denoised = lerp(uncond, cond, cfg)."""
    b = x.shape[0]
    out = model(torch.cat([x, x]), torch.cat([sigma, sigma]), torch.cat([uncond.expand(b, -1, -1), cond.expand(b, -1, -1)]))
    un, co = out.chunk(2)
    return torch.lerp(un, co, cfg)


def ancestral_step(sigma_from: Tensor, sigma_to: Tensor, eta: float = 1.0):
    """sampling_util.py:128-151."""
    sigma_up = min(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
    return sigma_down, sigma_up


def sample_euler_ancestral(denoise: Callable[[Tensor, Tensor], Tensor], x: Tensor, sigmas: Tensor,
                           noise_fn: Callable[[Tensor], Tensor]) -> Tensor:
    """sample_euler_ancestral_dy_cfg_pp as executed (samplers.py:612-740; SURVEY fact 7: the CFG++ branch is
    never taken, gamma = 0): plain Euler ancestral."""
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        den = denoise(x, sigmas[i] * s_in)
        sigma_down, sigma_up = ancestral_step(sigmas[i], sigmas[i + 1])
        d = (x - den) / sigmas[i]
        x = x + d * (sigma_down - sigmas[i])
        if sigmas[i + 1] > 0:
            x = x + noise_fn(x) * 1.0 * sigma_up
    return x


def multiscale_fullres(step: int, n_steps: int, start: int = 5, end: int = 8, intermittent: bool = True) -> bool:
    """should_use_fullres, samplers.py:838-857 with the sampler's own defaults (samplers.py:768-773)."""
    if step < start or step >= n_steps - end:
        return True
    if intermittent:
        return (step - start) % 2 == 0
    return False


def sample_dpmpp_2m(denoise: Callable[[Tensor, Tensor], Tensor], x: Tensor, sigmas: Tensor,
                    multiscale: bool = True, factor: float = 0.5, start: int = 5, end: int = 8,
                    intermittent: bool = True) -> Tensor:
    """sample_dpmpp_2m_cfgpp as executed (samplers.py:755-962): first-order exponential integrator
    x <- (sigma_{i+1}/sigma_i) x - expm1(-h) denoised, with the log->exp round trip of the reference, and the
    default-on multiscale schedule: low-res steps run the model on a bilinearly down-sampled x."""
    oh, ow = x.shape[-2:]
    sh = int(max(8, ((oh * factor) // 8) * 8)) if multiscale else oh
    sw = int(max(8, ((ow * factor) // 8) * 8)) if multiscale else ow
    active = multiscale and (sh != oh or sw != ow)
    t = -torch.log(sigmas)
    sigma_steps = torch.exp(-t)
    ratios = sigma_steps[1:] / sigma_steps[:-1]
    h = t[1:] - t[:-1]
    n = len(sigmas) - 1
    s_in = x.new_ones([x.shape[0]])
    for i in range(n):
        full = (not active) or multiscale_fullres(i, n, start, end, intermittent)
        xp = x if full else F.interpolate(x, size=(sh, sw), mode="bilinear", align_corners=False)
        den = denoise(xp, sigmas[i] * s_in)
        if not full:
            den = F.interpolate(den, size=(oh, ow), mode="bilinear", align_corners=False)
        x = ratios[i] * x - torch.expm1(-h[i]) * den
    return x


def prepare_noise(shape, seed: int) -> Tensor:
    """ksampler_util.py:274-295: global CPU generator seeded, one randn for the whole batch."""
    g = torch.manual_seed(seed)
    return torch.randn(shape, dtype=torch.float32, generator=g, device="cpu")


def ksample(sd, seed: int, steps: int, cfg: float, sampler: str, scheduler: str, cond: Tensor, uncond: Tensor,
            latent: Tensor, multiscale: bool = True, ms_options: Optional[dict] = None) -> Tensor:
    """KSampler.sample -> common_ksampler -> CFGGuider.sample -> KSAMPLER.sample for denoise=1.0 and an empty
    latent (sampling.py:773-887,1142-1233,445-497; CFG.py:266-294): returns samples / 0.18215."""
    tables = make_sigma_tables()
    sigmas = calculate_sigmas(scheduler, steps)
    noise = prepare_noise(latent.shape, seed)
    lat = latent * LATENT_SCALE if torch.count_nonzero(latent) > 0 else latent  # CFG.py:266-269
    max_denoise = math.isclose(float(tables[0][-1]), float(sigmas[0]), rel_tol=1e-05) or float(sigmas[0]) > float(tables[0][-1])
    x = noise * torch.sqrt(1.0 + sigmas[0] ** 2.0) if max_denoise else noise * sigmas[0]
    x = x + lat

    def model(xx, ss, cc):
        return apply_model(sd, xx, ss, cc, tables)

    def denoise(xx, ss):
        return cfg_denoise(model, xx, ss, cond, uncond, cfg)

    if sampler == "euler_ancestral_cfgpp":
        x = sample_euler_ancestral(denoise, x, sigmas, lambda t: torch.randn_like(t))
    elif sampler == "dpmpp_2m_cfgpp":
        x = sample_dpmpp_2m(denoise, x, sigmas, multiscale=multiscale, **(ms_options or {}))
    elif sampler == "euler_cfgpp":
        def denoise_pair(xx, ss):
            b = xx.shape[0]
            o = model(torch.cat([xx, xx]), torch.cat([ss, ss]), torch.cat([uncond.expand(b, -1, -1), cond.expand(b, -1, -1)]))
            return o[:b], o[b:]
        x = sample_euler_cfgpp(denoise_pair, x, sigmas, cfg)
    else:
        raise ValueError(sampler)
    return x / LATENT_SCALE


def sample_euler_cfgpp(denoise_pair: Callable[[Tensor, Tensor], Tuple[Tensor, Tensor]], x: Tensor, sigmas: Tensor, cfg: float,
                       cfg_scale: float = 7.5, cfg_min: float = 1.0) -> Tensor:
    """sample_euler_dy_cfg_pp AS EXECUTED (samplers.py:470-608). The sampler's own `uncond_denoised` bookkeeping is reset to
    None on every step by its manual post_cfg_function call (:548-550), so the main update is a plain Euler step on the
    guider's CFG result; what remains of "CFG++" is the extra dynamic step (dy_sampling_step_cfg_pp, :362-466) run while
    i // 2 == 1: the (1,1) pixel of every 2x2 block is denoised once more at HALF resolution -- at sigma_i although x has
    already been stepped to sigma_{i+1} -- with the guider's result extrapolated AGAIN from the true uncond by the sampler's
    own schedule current_cfg = cfg_scale + (cfg_min - cfg_scale) i / n (cfg_scale = 7.5 by default, not the user's cfg)."""
    n = len(sigmas) - 1
    s_in = x.new_ones([x.shape[0]])
    for i in range(n):
        current_cfg = cfg_scale + (cfg_min - cfg_scale) * (i / n)
        u, c = denoise_pair(x, sigmas[i] * s_in)
        den = u + (c - u) * cfg
        x = x + (x - den) / sigmas[i] * (sigmas[i + 1] - sigmas[i])
        if sigmas[i + 1] > 0 and i // 2 == 1:
            B, C, hh, ww = x.shape
            m, k = hh // 2, ww // 2
            sub = x[:, :, 1:2 * m:2, 1:2 * k:2].clone()  # a_list[:, :, :, 1, 1] of the 2x2 unfold
            u2, c2 = denoise_pair(sub, sigmas[i] * s_in)
            den2 = u2 + (c2 - u2) * cfg
            den2 = u2 + (den2 - u2) * current_cfg
            sub = sub + (sub - den2) / sigmas[i] * (sigmas[i + 1] - sigmas[i])
            x = x.clone()
            x[:, :, 1:2 * m:2, 1:2 * k:2] = sub
    return x


# ----------------------------------------------------------------------------------------------------------------
# VAE decoder  (AutoencodingEngine.decode / Decoder.forward, src/AutoEncoders/VariationalAE.py:130-145, 532-567;
# ResnetBlock src/AutoEncoders/ResBlock.py:383-406; AttnBlock src/Attention/Attention.py:159-178; Upsample
# VariationalAE.py:192-221; VAE.decode output mapping :602-604, 690-722)
# ----------------------------------------------------------------------------------------------------------------
VAE_CFG = dict(ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4, out_ch=3)
FLUX_VAE_CFG = dict(ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=16, out_ch=3)  # no (post_)quant_conv


def vae_decoder_param_shapes(cfg=VAE_CFG) -> Dict[str, Tuple[int, ...]]:
    """Decoder-side state-dict keys (without the 'first_stage_model.' prefix) -> shapes."""
    s: Dict[str, Tuple[int, ...]] = {}

    def conv(p, o, i, k):
        s[p + ".weight"] = (o, i, k, k)
        s[p + ".bias"] = (o,)

    def norm(p, c):
        s[p + ".weight"] = (c,)
        s[p + ".bias"] = (c,)

    def res(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cout, cin, 3)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".nin_shortcut", cout, cin, 1)

    ch, mult, nres = cfg["ch"], cfg["ch_mult"], cfg["num_res_blocks"]
    if cfg["z_channels"] == 4:  # the 16-channel Flux VAE has no post_quant_conv
        conv("post_quant_conv", cfg["z_channels"], cfg["z_channels"], 1)
    block_in = ch * mult[-1]
    conv("decoder.conv_in", block_in, cfg["z_channels"], 3)
    res("decoder.mid.block_1", block_in, block_in)
    norm("decoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"decoder.mid.attn_1.{n}", block_in, block_in, 1)
    res("decoder.mid.block_2", block_in, block_in)
    for lvl in reversed(range(len(mult))):
        block_out = ch * mult[lvl]
        for i in range(nres + 1):
            res(f"decoder.up.{lvl}.block.{i}", block_in, block_out)
            block_in = block_out
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", block_in, block_in, 3)
    norm("decoder.norm_out", block_in)
    conv("decoder.conv_out", cfg["out_ch"], block_in, 3)
    return s


def _vae_res(sd, p, x):
    h = _conv(sd, p + ".conv1", F.silu(_gn(sd, p + ".norm1", x, 1e-6)))
    h = _conv(sd, p + ".conv2", F.silu(_gn(sd, p + ".norm2", h, 1e-6)))
    if (p + ".nin_shortcut.weight") in sd:
        x = _conv(sd, p + ".nin_shortcut", x, padding=0)
    return x + h


def vae_decode(sd: Dict[str, Tensor], z: Tensor, cfg=VAE_CFG) -> Tensor:
    """z [B,4,h,w] fp32 -> image [B,8h,8w,3] fp32 in [0,1] (what VAEDecode returns)."""
    mult, nres = cfg["ch_mult"], cfg["num_res_blocks"]
    h = z.float()
    if "post_quant_conv.weight" in sd:  # absent in the Flux VAE (AutoencodingEngine flux=True, VariationalAE.py:103-145)
        h = _conv(sd, "post_quant_conv", h, padding=0)
    h = _conv(sd, "decoder.conv_in", h)
    h = _vae_res(sd, "decoder.mid.block_1", h)
    # AttnBlock: single head over all pixels, d = C
    p = "decoder.mid.attn_1"
    n = _gn(sd, p + ".norm", h, 1e-6)
    q, k, v = (_conv(sd, f"{p}.{t}", n, padding=0) for t in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    q, k, v = (t.reshape(b, c, hh * ww).transpose(1, 2) for t in (q, k, v))
    a = attention(q, k, v, heads=1).transpose(1, 2).reshape(b, c, hh, ww)
    h = h + _conv(sd, p + ".proj_out", a, padding=0)
    h = _vae_res(sd, "decoder.mid.block_2", h)
    for lvl in reversed(range(len(mult))):
        for i in range(nres + 1):
            h = _vae_res(sd, f"decoder.up.{lvl}.block.{i}", h)
        if lvl != 0:
            h = _conv(sd, f"decoder.up.{lvl}.upsample.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"))
    h = _conv(sd, "decoder.conv_out", F.silu(_gn(sd, "decoder.norm_out", h, 1e-6)))
    return torch.clamp((h + 1.0) / 2.0, min=0.0, max=1.0).movedim(1, -1)


def vae_encoder_param_shapes(cfg=VAE_CFG) -> Dict[str, Tuple[int, ...]]:
    """Encoder-side state-dict keys (without the 'first_stage_model.' prefix) -> shapes
    (Encoder.__init__ src/AutoEncoders/VariationalAE.py:260-356; quant_conv AutoencodingEngine :115-126)."""
    s: Dict[str, Tuple[int, ...]] = {}

    def conv(p, o, i, k):
        s[p + ".weight"] = (o, i, k, k)
        s[p + ".bias"] = (o,)

    def norm(p, c):
        s[p + ".weight"] = (c,)
        s[p + ".bias"] = (c,)

    def res(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cout, cin, 3)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".nin_shortcut", cout, cin, 1)

    ch, mult, nres, zc = cfg["ch"], cfg["ch_mult"], cfg["num_res_blocks"], cfg["z_channels"]
    conv("encoder.conv_in", ch, 3, 3)
    block_in = ch
    for lvl in range(len(mult)):
        block_out = ch * mult[lvl]
        for i in range(nres):
            res(f"encoder.down.{lvl}.block.{i}", block_in, block_out)
            block_in = block_out
        if lvl != len(mult) - 1:
            conv(f"encoder.down.{lvl}.downsample.conv", block_in, block_in, 3)
    res("encoder.mid.block_1", block_in, block_in)
    norm("encoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"encoder.mid.attn_1.{n}", block_in, block_in, 1)
    res("encoder.mid.block_2", block_in, block_in)
    norm("encoder.norm_out", block_in)
    conv("encoder.conv_out", 2 * zc, block_in, 3)
    conv("quant_conv", 2 * zc, 2 * zc, 1)
    return s


def vae_encode_moments(sd: Dict[str, Tensor], pixels: Tensor, cfg=VAE_CFG) -> Tensor:
    """pixels [B,H,W,3] fp32 in [0,1] -> Gaussian moments [B,8,H/8,W/8] fp32 (mean | logvar), i.e. quant_conv(Encoder(2x-1)).
    VAE.encode (VariationalAE.py:725-760) moves channels first and applies process_input (2x-1). NOTE: the reference's
    vae_encode_crop_pixels (:677-688) computes the cropped sizes and discards them -- pixels are NOT cropped; odd sizes
    are absorbed by the (0,1,0,1) pad + stride-2 convs (floor division at every level);
    Encoder.forward :377-413; Downsample pads (0,1,0,1) then conv stride 2 pad 0 (:224-254);
    AutoencodingEngine.encode :148-172 applies quant_conv, then DiagonalGaussianRegularizer samples
    mean + exp(0.5*clamp(logvar,-30,20)) * randn (:70-100) -- the stochastic part stays on the host."""
    mult, nres = cfg["ch_mult"], cfg["num_res_blocks"]
    B, H, W, _ = pixels.shape
    x = pixels[..., :3].movedim(-1, 1).float() * 2.0 - 1.0
    h = _conv(sd, "encoder.conv_in", x)
    for lvl in range(len(mult)):
        for i in range(nres):
            h = _vae_res(sd, f"encoder.down.{lvl}.block.{i}", h)
        if lvl != len(mult) - 1:
            h = _conv(sd, f"encoder.down.{lvl}.downsample.conv", F.pad(h, (0, 1, 0, 1)), stride=2, padding=0)
    h = _vae_res(sd, "encoder.mid.block_1", h)
    p = "encoder.mid.attn_1"
    n = _gn(sd, p + ".norm", h, 1e-6)
    q, k, v = (_conv(sd, f"{p}.{t}", n, padding=0) for t in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    q, k, v = (t.reshape(b, c, hh * ww).transpose(1, 2) for t in (q, k, v))
    a = attention(q, k, v, heads=1).transpose(1, 2).reshape(b, c, hh, ww)
    h = h + _conv(sd, p + ".proj_out", a, padding=0)
    h = _vae_res(sd, "encoder.mid.block_2", h)
    h = _conv(sd, "encoder.conv_out", F.silu(_gn(sd, "encoder.norm_out", h, 1e-6)))
    return _conv(sd, "quant_conv", h, padding=0)


def vae_sample_posterior(moments: Tensor, noise: Tensor) -> Tensor:
    """DiagonalGaussianDistribution.sample (VariationalAE.py:28-67): mean + exp(0.5 * clamp(logvar, -30, 20)) * noise."""
    mean, logvar = moments.chunk(2, dim=1)
    return mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * noise


def taesd_decoder_param_shapes() -> Dict[str, Tuple[int, ...]]:
    """Keys of taesd_decoder.safetensors = nn.Sequential indices of Decoder2 (src/AutoEncoders/taesd.py:104-136)."""
    s: Dict[str, Tuple[int, ...]] = {"1.weight": (64, 4, 3, 3), "1.bias": (64,)}
    for blk in (3, 4, 5, 8, 9, 10, 13, 14, 15, 18):
        for c in (0, 2, 4):
            s[f"{blk}.conv.{c}.weight"] = (64, 64, 3, 3)
            s[f"{blk}.conv.{c}.bias"] = (64,)
    for up in (7, 12, 17):
        s[f"{up}.weight"] = (64, 64, 3, 3)
    s["19.weight"] = (3, 64, 3, 3)
    s["19.bias"] = (3,)
    return s


def taesd_decode(sd: Dict[str, Tensor], z: Tensor) -> Tensor:
    """Raw latent [B,4,h,w] -> Decoder2(z) as [B,8h,8w,3] fp32 (taesd.py: Clamp :30-36, Block :39-63, Decoder2 :104-136).
    TAESD.decode (:190-197) returns this * 2 - 1; taesd_preview (:219-255) maps it back to [0,1] and clamps."""
    def blk(i, x):
        h = F.relu(_conv(sd, f"{i}.conv.0", x))
        h = F.relu(_conv(sd, f"{i}.conv.2", h))
        return F.relu(_conv(sd, f"{i}.conv.4", h) + x)

    def conv_nb(i, x):
        return F.conv2d(x, _w(sd, f"{i}.weight"), None, padding=1)

    x = torch.tanh(z.float() / 3) * 3
    x = F.relu(_conv(sd, "1", x))
    for i in (3, 4, 5):
        x = blk(i, x)
    x = conv_nb(7, F.interpolate(x, scale_factor=2.0, mode="nearest"))
    for i in (8, 9, 10):
        x = blk(i, x)
    x = conv_nb(12, F.interpolate(x, scale_factor=2.0, mode="nearest"))
    for i in (13, 14, 15):
        x = blk(i, x)
    x = conv_nb(17, F.interpolate(x, scale_factor=2.0, mode="nearest"))
    x = blk(18, x)
    return _conv(sd, "19", x).movedim(1, -1)


# ----------------------------------------------------------------------------------------------------------------
# CLIP-L text encoder  (CLIPTextModel_.forward src/clip/CLIPTextModel.py:51-107; CLIPLayer / CLIPAttention / CLIPMLP
# src/clip/Clip.py:14-180; config include/clip/sd1_clip_config.json: 12 layers, 768 wide, 12 heads, quick_gelu)
# ----------------------------------------------------------------------------------------------------------------
CLIP_CFG = dict(layers=12, width=768, heads=12, mlp=3072, vocab=49408, positions=77)


def clip_param_shapes(cfg=CLIP_CFG) -> Dict[str, Tuple[int, ...]]:
    """Keys below 'text_model.' -> shapes."""
    w, m = cfg["width"], cfg["mlp"]
    s: Dict[str, Tuple[int, ...]] = {"embeddings.token_embedding.weight": (cfg["vocab"], w),
                                     "embeddings.position_embedding.weight": (cfg["positions"], w)}
    for i in range(cfg["layers"]):
        p = f"encoder.layers.{i}"
        for n in ("layer_norm1", "layer_norm2"):
            s[f"{p}.{n}.weight"] = (w,)
            s[f"{p}.{n}.bias"] = (w,)
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s[f"{p}.self_attn.{n}.weight"] = (w, w)
            s[f"{p}.self_attn.{n}.bias"] = (w,)
        s[f"{p}.mlp.fc1.weight"] = (m, w)
        s[f"{p}.mlp.fc1.bias"] = (m,)
        s[f"{p}.mlp.fc2.weight"] = (w, m)
        s[f"{p}.mlp.fc2.bias"] = (w,)
    s["final_layer_norm.weight"] = (w,)
    s["final_layer_norm.bias"] = (w,)
    return s


def clip_encode(sd: Dict[str, Tensor], ids: Tensor, cfg=CLIP_CFG) -> Tuple[Tensor, Tensor]:
    """ids [S,77] int64 -> (final_LN(layer -2 output), final_LN(last layer output)), both [S,77,768] fp32.
    SD1.5 conditions on the first (CLIPSetLastLayer(-2), src/clip/Clip.py:592-608)."""
    heads = cfg["heads"]
    x = sd["embeddings.token_embedding.weight"].float()[ids] + sd["embeddings.position_embedding.weight"].float()
    n = x.shape[1]
    mask = torch.full((n, n), float("-inf")).triu_(1)
    inter = None
    for i in range(cfg["layers"]):
        p = f"encoder.layers.{i}"
        h = _ln(sd, p + ".layer_norm1", x)
        q = _linear(sd, p + ".self_attn.q_proj", h)
        k = _linear(sd, p + ".self_attn.k_proj", h)
        v = _linear(sd, p + ".self_attn.v_proj", h)
        x = x + _linear(sd, p + ".self_attn.out_proj", attention(q, k, v, heads, mask))
        h = _linear(sd, p + ".mlp.fc1", _ln(sd, p + ".layer_norm2", x))
        x = x + _linear(sd, p + ".mlp.fc2", h * torch.sigmoid(1.702 * h))
        if i == cfg["layers"] - 2:
            inter = x.clone()
    return _ln(sd, "final_layer_norm", inter), _ln(sd, "final_layer_norm", x)
