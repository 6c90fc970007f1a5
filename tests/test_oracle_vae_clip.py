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


def test_textual_inversion_rows_match_reference():
    """Token rows carrying textual-inversion vectors (the reference pipeline's default negative prompt has four): the
    product's host handling (pipeline.resolve_textual_embeddings / extend_token_table, Pipeline.encode) with the device call
    answered by the oracle reproduces the reference's SD1ClipModel.encode_token_weights -- extra ids past the vocabulary,
    a wrong-width vector dropped with the row re-padded and the weights left in place, vectors rounded to the table's dtype."""
    from lightdiffusion_next_b200.pipeline import Pipeline, extend_token_table, resolve_textual_embeddings
    g = torch.load(os.path.join(GOLDEN, "clip_ti_small.pt"))
    sd = dict(O.synth_state_dict(O.clip_param_shapes(), seed=777))
    row = [((g["vectors"][t[1]] if t[1] != "bad" else g["bad"]) if isinstance(t, tuple) else t, w) for t, w in g["row_spec"]]
    ids, wts, extra = resolve_textual_embeddings([row], 49408)
    assert ids.shape == (1, 77) and ids[0, :6].tolist() == [49406, 320, 49408, 49409, 1125, 49410] and int(ids[0, -1]) == 49407
    assert len(extra) == 3 and abs(float(wts[0, 3]) - 1.2) < 1e-6 and abs(float(wts[0, 6]) - 0.8) < 1e-6

    class Stand:
        device = torch.device("cpu")

        def __init__(self):
            self.sd = dict(sd)
            self.uploads = 0

        def clip_vocab(self):
            return 49408

        def set_clip_extra_embeddings(self, vectors):
            self.sd["embeddings.token_embedding.weight"] = extend_token_table(sd["embeddings.token_embedding.weight"], vectors)
            self.uploads += 1

        def clip_encode(self, ids):
            return O.clip_encode(self.sd, ids)

    e = Stand()
    cond = Pipeline(e).encode([row])
    assert e.uploads == 1 and e.sd["embeddings.token_embedding.weight"].shape == (49411, 768)
    assert rel(cond, g["cond"]) < 2e-5
    plain = Pipeline(e).encode([[(49406, 1.0), (320, 1.0)] + [(49407, 1.0)] * 75])     # the original rows still work
    assert rel(plain, O.clip_encode(sd, torch.tensor([[49406, 320] + [49407] * 75]))[0]) < 1e-6
