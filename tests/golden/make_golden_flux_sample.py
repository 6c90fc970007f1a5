"""Generates tests/golden/flux_sample_small.pt: full `KSampler.sample(..., sampler_name="euler_cfgpp", scheduler="beta",
flux=True)` runs of the UNMODIFIED reference (src/sample/sampling.py:773-887 -> CFGGuider(flux=True) -> ModelSamplingFlux +
CONST -> sample_euler_dy_cfg_pp) around its own Flux2/Flux3 model (src/BlackForest/Flux.py:543-843) on a small
configuration with seeded synthetic weights -- the Flux branch of pipeline() (src/user/pipeline.py:251-264) end to end
minus the text encoders and the VAE (build container only)."""
import os
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.modules.setdefault("torchsde", types.ModuleType("torchsde"))
os.chdir(tempfile.mkdtemp(prefix="ldn_golden_"))
from oracle import flux_oracle as FO  # noqa: E402
from oracle import sd15_oracle as O  # noqa: E402

torch.manual_seed(0)
torch.set_grad_enabled(False)
from src.user import app_instance  # noqa: E402
app_instance.app.previewer_var.set(False)
from src.BlackForest import Flux as RF  # noqa: E402
from src.cond import cast  # noqa: E402
from src.Model import ModelPatcher  # noqa: E402
from src.sample import sampling  # noqa: E402

cfg = dict(FO.FLUX_TINY)
mc = RF.Flux(dict(cfg, image_model="flux"))
mc.custom_operations = cast.disable_weight_init
mc.set_inference_dtype(torch.float32, None)
model = mc.get_model({}, "", device=torch.device("cpu"))
gold = torch.load(os.path.join(HERE, "flux_small.pt"))
sd = {k: v.float() for k, v in O.synth_state_dict(FO.flux_param_shapes(cfg), seed=8642).items()}
for k, v in gold["scales"].items():
    sd[k] = v.float()
model.diffusion_model.load_state_dict(sd, strict=True)
mp = ModelPatcher.ModelPatcher(model, load_device=torch.device("cpu"), offload_device=torch.device("cpu"))
g = torch.Generator().manual_seed(77)
nt = 24
t5_pos = torch.randn(1, nt, cfg["context_in_dim"], generator=g)
y_pos = torch.randn(1, cfg["vec_in_dim"], generator=g)
out = {"t5_pos": t5_pos, "y_pos": y_pos}
# negative = ConditioningZeroOut of the positive (pipeline.py:247-249): zero states, zero pooled vector
for name, steps, h, w, cfg_scale, guidance, zero_neg, batch in (("a", 6, 16, 16, 1.0, 3.0, True, 1), ("b", 5, 12, 20, 2.5, 3.5, False, 1),
                                                             ("c", 4, 8, 12, 1.0, 3.0, True, 2)):
    if zero_neg:
        t5_neg, y_neg = torch.zeros_like(t5_pos), torch.zeros_like(y_pos)
    else:
        t5_neg, y_neg = torch.randn(1, nt, cfg["context_in_dim"], generator=g), torch.randn(1, cfg["vec_in_dim"], generator=g)
    res = sampling.KSampler().sample(
        model=mp.clone(), seed=42, steps=steps, cfg=cfg_scale, sampler_name="euler_cfgpp", scheduler="beta", denoise=1.0,
        positive=[[t5_pos, {"pooled_output": y_pos, "guidance": guidance}]],
        negative=[[t5_neg, {"pooled_output": y_neg, "guidance": guidance}]],
        latent_image={"samples": torch.zeros(batch, 16, h, w)}, pipeline=True, flux=True)
    out[f"{name}_final"] = res[0]["samples"].clone()
    out[f"{name}_args"] = dict(steps=steps, h=h, w=w, cfg=cfg_scale, guidance=guidance, batch=batch)
    out[f"{name}_t5_neg"], out[f"{name}_y_neg"] = t5_neg, y_neg
    print(name, tuple(res[0]["samples"].shape), float(res[0]["samples"].mean()), float(res[0]["samples"].std()), flush=True)
torch.save(out, os.path.join(HERE, "flux_sample_small.pt"))
print("wrote flux_sample_small.pt")
