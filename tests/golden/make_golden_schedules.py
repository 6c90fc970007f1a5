"""Generates tests/golden/schedules.pt: sigma schedules of every scheduler name the reference's calculate_sigmas accepts
(src/sample/ksampler_util.py:244-271), from the UNMODIFIED reference (build container only)."""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.modules.setdefault("torchsde", types.ModuleType("torchsde"))
import tempfile
os.chdir(tempfile.mkdtemp(prefix="ldn_golden_"))
from src.sample import ksampler_util, sampling  # noqa: E402

ms = sampling.ModelSamplingDiscrete(types.SimpleNamespace(sampling_settings={}))
out = {}
for name in ("karras", "normal", "simple", "beta"):
    for steps in (1, 4, 10, 20, 30, 50):
        out[f"{name}_{steps}"] = ksampler_util.calculate_sigmas(ms, name, steps).clone()
        print(name, steps, tuple(out[f"{name}_{steps}"].shape))
torch.save(out, os.path.join(HERE, "schedules.pt"))
