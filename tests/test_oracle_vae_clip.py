"""CPU: pins the oracle's VAE decoder and CLIP-L encoder against fixtures produced by the reference itself
(tests/golden/make_golden.py vae clip)."""
import os

import torch

from oracle import sd15_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def test_vae_decode_matches_reference():
    g = torch.load(os.path.join(GOLDEN, "vae_small.pt"))
    sd = O.synth_state_dict(O.vae_decoder_param_shapes(), seed=4321)  # fp16-rounded, like the fixture
    for hw in (8, 16):
        img = O.vae_decode(sd, g[f"z_{hw}"])
        assert img.shape == g[f"img_{hw}"].shape == (2, 8 * hw, 8 * hw, 3)
        assert rel(img, g[f"img_{hw}"]) < 1e-5
        assert float(img.min()) >= 0.0 and float(img.max()) <= 1.0


def test_clip_encode_matches_reference():
    g = torch.load(os.path.join(GOLDEN, "clip_small.pt"))
    sd = O.synth_state_dict(O.clip_param_shapes(), seed=777)
    pen_empty, _ = O.clip_encode(sd, g["empty_ids"])
    for name in ("plain", "empty"):
        pen, last = O.clip_encode(sd, g[f"{name}_ids"])
        assert rel(pen, g[f"{name}_cond"]) < 2e-5, name      # SD1.5 conditions on layer -2 (+ final LN)
        assert rel(last, g[f"{name}_cond"]) > 1e-2
    # prompt weighting (ClipTokenWeightEncoder, src/SD15/SDClip.py:54-76): z = (z - z_empty) * w + z_empty per token
    pen, _ = O.clip_encode(sd, g["weighted_ids"])
    w = g["weighted_weights"][0][:, None]
    z = (pen[0] - pen_empty[0]) * w + pen_empty[0]
    assert rel(z[None], g["weighted_cond"]) < 2e-5
    assert float((g["weighted_weights"] != 1).sum()) >= 2


def test_oracle_vae_encoder_matches_reference_golden():
    """oracle.vae_encode_moments / vae_sample_posterior vs the reference Encoder + quant_conv + regulariser
    (fixture: tests/golden/make_golden_vae_enc.py; b has a width that is not a multiple of 8)."""
    import os
    import torch
    from oracle import sd15_oracle as O
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "vae_enc_small.pt"))
    sd = O.synth_state_dict(O.vae_encoder_param_shapes(), seed=2468)
    for name in ("a", "b"):
        m = O.vae_encode_moments(sd, gold[f"pixels_{name}"])
        ref = gold[f"moments_{name}"]
        assert m.shape == ref.shape
        assert ((m - ref).norm() / ref.norm()).item() < 1e-4
        torch.manual_seed(int(gold["sample_seed"]))
        lat = O.vae_sample_posterior(m, torch.randn(ref[:, :4].shape))
        assert ((lat - gold[f"latent_{name}"]).norm() / gold[f"latent_{name}"].norm()).item() < 1e-4


def test_oracle_taesd_decoder_matches_reference_golden():
    """oracle.taesd_decode vs the reference's Decoder2 (fixture: tests/golden/make_golden_taesd.py)."""
    import os
    import torch
    from oracle import sd15_oracle as O
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "taesd_small.pt"))
    sd = O.synth_state_dict(O.taesd_decoder_param_shapes(), seed=1357)
    for name in ("a", "b"):
        y = O.taesd_decode(sd, gold[f"z_{name}"])
        ref = gold[f"dec_{name}"]
        assert y.shape == ref.shape
        assert ((y - ref).norm() / ref.norm()).item() < 1e-5


def test_oracle_flux_vae_decoder_matches_reference_golden():
    """16-channel Flux VAE decoder (no post_quant_conv): oracle vs the reference's AutoencodingEngine(flux=True)."""
    import os
    import torch
    from oracle import sd15_oracle as O
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "flux_vae_small.pt"))
    sd = O.synth_state_dict(O.vae_decoder_param_shapes(O.FLUX_VAE_CFG), seed=9753)
    assert "post_quant_conv.weight" not in sd
    for name in ("a", "b"):
        img = O.vae_decode(sd, gold[f"z_{name}"])
        ref = gold[f"img_{name}"]
        assert img.shape == ref.shape and ((img - ref).norm() / ref.norm()).item() < 1e-4
