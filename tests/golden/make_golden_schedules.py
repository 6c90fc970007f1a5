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
fm = sampling.ModelSamplingFlux()
out["flux_sigmas"] = fm.sigmas.clone()
for name in ("simple", "beta"):  # ModelSamplingFlux has no sigma_min: karras / normal raise in the reference itself
    for steps in (4, 20, 28):
        out[f"flux__{name}_{steps}"] = ksampler_util.calculate_sigmas(fm, name, steps).clone()
torch.save(out, os.path.join(HERE, "schedules.pt"))
