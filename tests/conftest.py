import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def unet_sd():
    """Seeded synthetic SD1.5 UNet weights (fp16, as the reference stores them) — oracle/sd15_oracle.py."""
    from oracle import sd15_oracle as O

    return O.synth_state_dict(O.unet_param_shapes())


# A narrow UNet of the SD1.5 topology (same blocks, attention levels, context width) for HOST-LOGIC tests that compare
# the product's composition with a by-hand composition over the same stand-in engine and pin nothing to a golden:
# the full-width oracle forward costs ~0.7 s on the host, this one a few ms.
TINY_UNET = dict(in_channels=4, out_channels=4, model_channels=32, channel_mult=(1, 2, 4, 4), num_res_blocks=2,
                 attn_levels=(True, True, True, False), num_heads=8, context_dim=768, time_embed_dim=128)


@pytest.fixture(scope="session")
def tiny_unet_sd():
    from oracle import sd15_oracle as O

    return O.synth_state_dict(O.unet_param_shapes(TINY_UNET), seed=99)


@pytest.fixture(scope="session")
def golden_unet():
    import torch

    return torch.load(os.path.join(GOLDEN, "unet_small.pt"))


@pytest.fixture(scope="session")
def golden_sample():
    import torch

    return torch.load(os.path.join(GOLDEN, "sample_small.pt"))
