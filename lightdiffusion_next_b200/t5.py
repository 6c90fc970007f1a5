"""Host side of the Flux text encoders: the T5-XXL branch of the reference's FluxClipModel (src/clip/FluxClip.py:675-718)
around `ldn_t5_encode`, and the CLIP-L pooled vector the Flux model takes as `y`.

What stays on the host, with the reference counterpart of each piece:
  relative_position_buckets  <- T5Attention._relative_position_bucket / compute_bias (FluxClip.py:153-240): one bucket per
                                relative distance, evaluated with the reference's own fp32 torch arithmetic so that the
                                bucket boundaries agree bit for bit (the device only gathers table[bucket, head])
  pad_tokens                 <- T5XXLTokenizer row format (FluxClip.py:593-613): ids + end (1), zero-padded to 256
  encode_token_weights       <- ClipTokenWeightEncoder.encode_token_weights (src/SD15/SDClip.py:33-97): weighted tokens are
                                pulled towards the empty prompt's states
No CPU fallback: the transformer stack itself only exists as the CUDA program (csrc/t5.cu).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch

NUM_BUCKETS = 32
MAX_DISTANCE = 128
MIN_LENGTH = 256
END_TOKEN, PAD_TOKEN = 1, 0


def t5_shapes(d_model: int = 4096, d_ff: int = 10240, num_heads: int = 64, num_layers: int = 24,
              vocab_size: int = 32128) -> Dict[str, Tuple[int, ...]]:
    """State-dict layout `Engine.load_t5` expects (keys of the reference's T5 module; defaults = t5_config_xxl.json)."""
    s: Dict[str, Tuple[int, ...]] = {}
    for i in range(num_layers):
        p = f"encoder.block.{i}.layer"
        for n in "qkvo":
            s[f"{p}.0.SelfAttention.{n}.weight"] = (d_model, d_model)
        if i == 0:
            s[f"{p}.0.SelfAttention.relative_attention_bias.weight"] = (NUM_BUCKETS, num_heads)
        s[f"{p}.0.layer_norm.weight"] = (d_model,)
        s[f"{p}.1.DenseReluDense.wi_0.weight"] = (d_ff, d_model)
        s[f"{p}.1.DenseReluDense.wi_1.weight"] = (d_ff, d_model)
        s[f"{p}.1.DenseReluDense.wo.weight"] = (d_model, d_ff)
        s[f"{p}.1.layer_norm.weight"] = (d_model,)
    s["encoder.final_layer_norm.weight"] = (d_model,)
    s["shared.weight"] = (vocab_size, d_model)
    return s


def validate_state_dict(sd) -> Dict[str, int]:
    """Checks a T5 encoder state dict against the layout the engine consumes (names and shapes, decoder / lm_head tensors
    tolerated and ignored by the caller) and returns its configuration.  Raises ValueError naming what is wrong -- the
    engine has no fallback for a malformed checkpoint."""
    if "shared.weight" not in sd or sd["shared.weight"].dim() != 2:
        raise ValueError("T5 state dict: shared.weight [vocab, d_model] is missing")
    vocab, d_model = (int(v) for v in sd["shared.weight"].shape)
    rb = "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"
    wi = "encoder.block.0.layer.1.DenseReluDense.wi_0.weight"
    if rb not in sd or wi not in sd:
        raise ValueError(f"T5 state dict: {rb if rb not in sd else wi} is missing (a gated-activation T5 v1.1 encoder is expected)")
    heads, d_ff = int(sd[rb].shape[1]), int(sd[wi].shape[0])
    layers = 0
    while f"encoder.block.{layers}.layer.0.SelfAttention.q.weight" in sd:
        layers += 1
    want = t5_shapes(d_model, d_ff, heads, layers, vocab)
    missing = sorted(k for k in want if k not in sd)
    wrong = sorted(k for k in want if k in sd and tuple(sd[k].shape) != want[k])
    if missing or wrong:
        raise ValueError(f"T5 state dict does not match the encoder layout: missing {missing[:4]}{'...' if len(missing) > 4 else ''}, "
                         f"wrong shape {[(k, tuple(sd[k].shape), want[k]) for k in wrong[:3]]}")
    if d_model % heads != 0 or d_model // heads != 64:
        raise ValueError(f"T5: head width {d_model / heads:g} is not supported (the engine's logit-bias attention is built for 64)")
    return dict(vocab_size=vocab, d_model=d_model, d_ff=d_ff, num_heads=heads, num_layers=layers)


def relative_position_buckets(n: int) -> torch.Tensor:
    """int32 [2n-1]: bucket of the relative distance (key - query) = -(n-1) .. n-1 (bidirectional, 32 buckets, distances
    >= 128 share the last bucket of their direction).  bias[h, i, j] = table[buckets[j - i + n - 1], h]."""
    rel = torch.arange(-(n - 1), n, dtype=torch.long)
    half = NUM_BUCKETS // 2
    out = (rel > 0).to(torch.long) * half
    rel = torch.abs(rel)
    max_exact = half // 2
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(MAX_DISTANCE / max_exact) * (half - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, half - 1))
    out = out + torch.where(rel < max_exact, rel, large)
    return out.to(torch.int32)


def pad_tokens(ids: Sequence[int], min_length: int = MIN_LENGTH) -> List[Tuple[int, float]]:
    """Sentencepiece ids of a prompt (without the end token) -> the (token, weight) row the reference's tokenizer emits."""
    row = [(int(t), 1.0) for t in ids] + [(END_TOKEN, 1.0)]
    row += [(PAD_TOKEN, 1.0)] * max(0, min_length - len(row))
    return row


def encode_token_weights(engine, token_weight_pairs: Sequence[Sequence[Tuple[int, float]]]) -> torch.Tensor:
    """[(token, weight)] sections of equal length -> T5 states [1, sections * n, d_model] fp32 on the engine's device."""
    sections = [[t for t, _ in sec] for sec in token_weight_pairs]
    if len(sections) == 0 or len({len(s) for s in sections}) != 1:
        raise ValueError("encode_token_weights: sections must be non-empty and of equal length")
    n = len(sections[0])
    has_weights = any(w != 1.0 for sec in token_weight_pairs for _, w in sec)
    rows = sections + ([[END_TOKEN] + [PAD_TOKEN] * (n - 1)] if has_weights else [])
    out = engine.t5_encode(torch.tensor(rows, dtype=torch.long))
    res = []
    for k, sec in enumerate(token_weight_pairs):
        z = out[k]
        if has_weights:
            z_empty = out[-1]
            w = torch.tensor([wt for _, wt in sec], dtype=torch.float32, device=z.device).view(-1, 1)
            z = torch.where(w != 1.0, (z - z_empty) * w + z_empty, z)
        res.append(z)
    return torch.cat(res, dim=0).unsqueeze(0)
