"""Shared set-up for the golden generators that import the UNMODIFIED reference (read-only, /root/reference).
Build container only -- the GPU box has no /root/reference; the fixtures these scripts write are committed.
Recipe = SURVEY.md Appendix C (torchsde stub, previewer off, cwd with ./include/clip and ./include/sd1_tokenizer)."""
import os
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

SD15_UNET_CONFIG = dict(image_size=32, in_channels=4, out_channels=4, model_channels=320, num_res_blocks=[2, 2, 2, 2],
                        channel_mult=[1, 2, 4, 4], transformer_depth=[1, 1, 1, 1, 1, 1, 0, 0],
                        transformer_depth_output=[1] * 9 + [0] * 3, transformer_depth_middle=1,
                        use_linear_in_transformer=False, context_dim=768, use_spatial_transformer=True, legacy=False,
                        use_checkpoint=False, adm_in_channels=None, use_temporal_attention=False,
                        use_temporal_resblock=False)


def enter_reference():
    """sys.path, torchsde stub, a scratch cwd laid out the way the reference expects; previewer off."""
    for p in (ROOT, REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.modules.setdefault("torchsde", types.ModuleType("torchsde"))
    work = tempfile.mkdtemp(prefix="ldn_golden_")
    os.makedirs(os.path.join(work, "include"), exist_ok=True)
    for sub in ("clip", "sd1_tokenizer"):
        os.symlink(os.path.join(REF, "include", sub), os.path.join(work, "include", sub))
    os.chdir(work)
    torch.set_grad_enabled(False)
    from src.user import app_instance
    app_instance.app.previewer_var.set(False)
    return work


def build_reference_unet(device="cpu"):
    """The reference's own BaseModel + ModelPatcher around UNetModel1 with the seeded synthetic weights of the oracle."""
    from oracle import sd15_oracle as O
    from src.Device import Device
    from src.Model import ModelPatcher
    from src.NeuralNetwork import unet

    mc = unet.model_config_from_unet_config(dict(SD15_UNET_CONFIG))
    dt = unet.unet_dtype1()
    mc.set_inference_dtype(dt, Device.unet_manual_cast(dt, Device.get_torch_device(), mc.supported_inference_dtypes))
    model = mc.get_model({}, "", device=torch.device(device))
    sd = O.synth_state_dict(O.unet_param_shapes())
    model.diffusion_model.load_state_dict(sd, strict=True)
    mp = ModelPatcher.ModelPatcher(model, load_device=torch.device(device), offload_device=torch.device(device))
    return model, mp, sd


def contexts(seed=1234):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(1, 77, 768, generator=g), torch.randn(1, 77, 768, generator=g)


class SeamRecorder:
    """model_function_wrapper seam (src/cond/cond.py:254-265): pass-through that records every call."""

    def __init__(self):
        self.calls = []

    def __call__(self, model_function, params):
        r = model_function(params["input"], params["timestep"], **params["c"])
        self.calls.append(dict(input=params["input"].clone(), timestep=params["timestep"].clone(), output=r.clone(),
                               ctx_id=id(params["c"]["c_crossattn"])))
        return r

    def to(self, *_):
        return self


class SeqNoise:
    """Deterministic stand-in for the Brownian tree: the n-th call returns the n-th draw of a seeded CPU generator."""

    def __init__(self, shape, seed):
        self.g = torch.Generator().manual_seed(seed)
        self.shape = tuple(shape)
        self.calls = []

    def __call__(self, sigma, sigma_next):
        self.calls.append((float(sigma), float(sigma_next)))
        return torch.randn(self.shape, generator=self.g)
