"""GPU: the pipe(prompt...) / sample() / decode() surface end to end on synthetic weights."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())


def test_pipeline_tokens_to_image(unet_sd):
    from lightdiffusion_next_b200.engine import Engine
    from lightdiffusion_next_b200.pipeline import EMPTY_TOKENS, Pipeline
    from oracle import sd15_oracle as O
    eng = Engine(max_rows=4, max_h=32, max_w=32)
    eng.load_unet(unet_sd)
    vsd = O.synth_state_dict(O.vae_decoder_param_shapes(), seed=4321)
    csd = O.synth_state_dict(O.clip_param_shapes(), seed=777)
    eng.load_vae(vsd)
    eng.load_clip(csd)
    pipe = Pipeline(eng)
    g = torch.load(os.path.join(GOLDEN, "clip_small.pt"))
    assert g["empty_ids"][0].tolist() == EMPTY_TOKENS
    # CLIPTextEncode with prompt weights == the reference's weighted conditioning
    toks = [list(zip(g["weighted_ids"][0].tolist(), g["weighted_weights"][0].tolist()))]
    cond = pipe.encode(toks)
    assert cond.shape == (1, 77, 768) and rel(cond, g["weighted_cond"]) < 2e-2
    plain = [[(t, 1.0) for t in g["plain_ids"][0].tolist()]]
    # pooled vector (Flux `y`): last layer + final LN at the first end-of-text token (CLIPTextModel.py:95-105)
    cond_p, pooled = pipe.encode(plain, return_pooled=True)
    _, last = O.clip_encode(csd, g["plain_ids"])
    eos = int((g["plain_ids"][0] == 49407).int().argmax())
    assert pooled.shape == (1, 768) and rel(pooled, last[0:1, eos]) < 2e-2
    # full path: 2 images, 3 steps, 256x256; compared with the oracle run on the same tokens / seed
    img = pipe(plain, None, width=256, height=256, batch=2, seed=7, steps=3, cfg=7.0)
    assert img.shape == (2, 256, 256, 3) and torch.isfinite(img).all()
    assert float(img.min()) >= 0 and float(img.max()) <= 1
    pos, _ = O.clip_encode(csd, g["plain_ids"])
    neg, _ = O.clip_encode(csd, g["empty_ids"])
    lat = O.ksample(unet_sd, 7, 3, 7.0, "dpmpp_2m_cfgpp", "karras", pos, neg, torch.zeros(2, 4, 32, 32))
    ref = O.vae_decode(vsd, lat)
    assert rel(img, ref) < 5e-2, rel(img, ref)
