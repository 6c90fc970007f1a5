"""Seeded synthetic SD1.5 weights in the reference's state-dict layout (LDM key names, fp16 storage).

No checkpoints exist offline, so benchmarks and smoke tests drive the engine with these.  Per-tensor seed =
crc32(name) ^ seed, so the values do not depend on iteration order and rank 0 can regenerate any tensor.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Tuple

import torch

MODEL_CH, TEMB, CTX, HEADS = 320, 1280, 768, 8
CH_MULT = (1, 2, 4, 4)
ATTN = (True, True, True, False)
NUM_RES = 2
_RES_OUT = ("out_layers.3.", "proj_out.", "to_out.0.", "ff.net.2.", "conv2.", "out_proj.", "mlp.fc2.")


def unet_shapes() -> Dict[str, Tuple[int, ...]]:
    s: Dict[str, Tuple[int, ...]] = {}

    def wb(p, *shape):
        s[p + ".weight"] = tuple(shape)
        s[p + ".bias"] = (shape[0],)

    def res(p, cin, cout):
        s[p + ".in_layers.0.weight"] = s[p + ".in_layers.0.bias"] = (cin,)
        wb(p + ".in_layers.2", cout, cin, 3, 3)
        wb(p + ".emb_layers.1", cout, TEMB)
        s[p + ".out_layers.0.weight"] = s[p + ".out_layers.0.bias"] = (cout,)
        wb(p + ".out_layers.3", cout, cout, 3, 3)
        if cin != cout:
            wb(p + ".skip_connection", cout, cin, 1, 1)

    def st(p, c):
        s[p + ".norm.weight"] = s[p + ".norm.bias"] = (c,)
        wb(p + ".proj_in", c, c, 1, 1)
        wb(p + ".proj_out", c, c, 1, 1)
        t = p + ".transformer_blocks.0"
        for a, kd in (("attn1", c), ("attn2", CTX)):
            s[f"{t}.{a}.to_q.weight"] = (c, c)
            s[f"{t}.{a}.to_k.weight"] = (c, kd)
            s[f"{t}.{a}.to_v.weight"] = (c, kd)
            wb(f"{t}.{a}.to_out.0", c, c)
        wb(f"{t}.ff.net.0.proj", 8 * c, c)
        wb(f"{t}.ff.net.2", c, 4 * c)
        for n in ("norm1", "norm2", "norm3"):
            s[f"{t}.{n}.weight"] = s[f"{t}.{n}.bias"] = (c,)

    wb("time_embed.0", TEMB, MODEL_CH)
    wb("time_embed.2", TEMB, TEMB)
    wb("input_blocks.0.0", MODEL_CH, 4, 3, 3)
    ch, idx, chans = MODEL_CH, 1, [MODEL_CH]
    for lvl, m in enumerate(CH_MULT):
        for _ in range(NUM_RES):
            res(f"input_blocks.{idx}.0", ch, m * MODEL_CH)
            ch = m * MODEL_CH
            if ATTN[lvl]:
                st(f"input_blocks.{idx}.1", ch)
            chans.append(ch)
            idx += 1
        if lvl < len(CH_MULT) - 1:
            wb(f"input_blocks.{idx}.0.op", ch, ch, 3, 3)
            chans.append(ch)
            idx += 1
    res("middle_block.0", ch, ch)
    st("middle_block.1", ch)
    res("middle_block.2", ch, ch)
    idx = 0
    for lvl in reversed(range(len(CH_MULT))):
        for i in range(NUM_RES + 1):
            res(f"output_blocks.{idx}.0", ch + chans.pop(), MODEL_CH * CH_MULT[lvl])
            ch = MODEL_CH * CH_MULT[lvl]
            sub = 1
            if ATTN[lvl]:
                st(f"output_blocks.{idx}.1", ch)
                sub = 2
            if lvl > 0 and i == NUM_RES:
                wb(f"output_blocks.{idx}.{sub}.conv", ch, ch, 3, 3)
            idx += 1
    s["out.0.weight"] = s["out.0.bias"] = (MODEL_CH,)
    wb("out.2", 4, MODEL_CH, 3, 3)
    return s


def synth_tensor(name: str, shape: Tuple[int, ...], seed: int = 1234, dtype=torch.float16) -> torch.Tensor:
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ seed) & 0x7FFFFFFF)
    if len(shape) > 1:
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        if "embedding" in name:
            w = torch.randn(shape, generator=g) * 0.02
        else:
            w = torch.randn(shape, generator=g) / math.sqrt(fan_in)
            if any(k in name for k in _RES_OUT):
                w = w * 0.5
    elif name.endswith(".weight"):
        w = 1.0 + 0.1 * torch.randn(shape, generator=g)
    else:
        w = 0.05 * torch.randn(shape, generator=g)
    return w.to(dtype)


def synth_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 1234, dtype=torch.float16) -> Dict[str, torch.Tensor]:
    return {k: synth_tensor(k, v, seed, dtype) for k, v in shapes.items()}


def vae_decoder_shapes(ch: int = 128, ch_mult=(1, 2, 4, 4), num_res: int = 2, z: int = 4, out_ch: int = 3) -> Dict[str, Tuple[int, ...]]:
    """State-dict layout of the SD1.5 VAE decoder side (keys below `first_stage_model.`; Decoder,
    src/AutoEncoders/VariationalAE.py:416-567)."""
    s: Dict[str, Tuple[int, ...]] = {}

    def conv(p, o, i, k):
        s[p + ".weight"] = (o, i, k, k)
        s[p + ".bias"] = (o,)

    def norm(p, c):
        s[p + ".weight"] = s[p + ".bias"] = (c,)

    def res(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cout, cin, 3)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".nin_shortcut", cout, cin, 1)

    conv("post_quant_conv", z, z, 1)
    c = ch * ch_mult[-1]
    conv("decoder.conv_in", c, z, 3)
    res("decoder.mid.block_1", c, c)
    norm("decoder.mid.attn_1.norm", c)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"decoder.mid.attn_1.{n}", c, c, 1)
    res("decoder.mid.block_2", c, c)
    for lvl in reversed(range(len(ch_mult))):
        co = ch * ch_mult[lvl]
        for i in range(num_res + 1):
            res(f"decoder.up.{lvl}.block.{i}", c, co)
            c = co
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", c, c, 3)
    norm("decoder.norm_out", c)
    conv("decoder.conv_out", out_ch, c, 3)
    return s


def vae_encoder_shapes(ch: int = 128, ch_mult=(1, 2, 4, 4), num_res: int = 2, z: int = 4) -> Dict[str, Tuple[int, ...]]:
    """State-dict layout of the SD1.5 VAE encoder side + quant_conv (Encoder, src/AutoEncoders/VariationalAE.py:257-413)."""
    s: Dict[str, Tuple[int, ...]] = {}

    def conv(p, o, i, k):
        s[p + ".weight"] = (o, i, k, k)
        s[p + ".bias"] = (o,)

    def norm(p, c):
        s[p + ".weight"] = s[p + ".bias"] = (c,)

    def res(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cout, cin, 3)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".nin_shortcut", cout, cin, 1)

    conv("encoder.conv_in", ch, 3, 3)
    c = ch
    for lvl in range(len(ch_mult)):
        co = ch * ch_mult[lvl]
        for i in range(num_res):
            res(f"encoder.down.{lvl}.block.{i}", c, co)
            c = co
        if lvl != len(ch_mult) - 1:
            conv(f"encoder.down.{lvl}.downsample.conv", c, c, 3)
    res("encoder.mid.block_1", c, c)
    norm("encoder.mid.attn_1.norm", c)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"encoder.mid.attn_1.{n}", c, c, 1)
    res("encoder.mid.block_2", c, c)
    norm("encoder.norm_out", c)
    conv("encoder.conv_out", 2 * z, c, 3)
    conv("quant_conv", 2 * z, 2 * z, 1)
    return s


def taesd_decoder_shapes(latent_channels: int = 4) -> Dict[str, Tuple[int, ...]]:
    """State-dict layout of `taesd_decoder.safetensors` (Decoder2, src/AutoEncoders/taesd.py:104-136): nn.Sequential indices."""
    s: Dict[str, Tuple[int, ...]] = {"1.weight": (64, latent_channels, 3, 3), "1.bias": (64,)}
    for blk in (3, 4, 5, 8, 9, 10, 13, 14, 15, 18):
        for c in (0, 2, 4):
            s[f"{blk}.conv.{c}.weight"] = (64, 64, 3, 3)
            s[f"{blk}.conv.{c}.bias"] = (64,)
    for up in (7, 12, 17):
        s[f"{up}.weight"] = (64, 64, 3, 3)  # bias=False
    s["19.weight"] = (3, 64, 3, 3)
    s["19.bias"] = (3,)
    return s


def clip_shapes(width: int = 768, mlp: int = 3072, layers: int = 12, vocab: int = 49408, positions: int = 77) -> Dict[str, Tuple[int, ...]]:
    """State-dict layout of CLIP-L's text model (keys below `text_model.`; CLIPTextModel_, src/clip/CLIPTextModel.py:3-107)."""
    s: Dict[str, Tuple[int, ...]] = {"embeddings.token_embedding.weight": (vocab, width),
                                     "embeddings.position_embedding.weight": (positions, width)}
    for i in range(layers):
        p = f"encoder.layers.{i}"
        for n in ("layer_norm1", "layer_norm2"):
            s[f"{p}.{n}.weight"] = s[f"{p}.{n}.bias"] = (width,)
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s[f"{p}.self_attn.{n}.weight"] = (width, width)
            s[f"{p}.self_attn.{n}.bias"] = (width,)
        s[f"{p}.mlp.fc1.weight"] = (mlp, width)
        s[f"{p}.mlp.fc1.bias"] = (mlp,)
        s[f"{p}.mlp.fc2.weight"] = (width, mlp)
        s[f"{p}.mlp.fc2.bias"] = (width,)
    s["final_layer_norm.weight"] = s["final_layer_norm.bias"] = (width,)
    return s
