"""ctypes binding of libldn.so (include/ldn.h).  No fallback: if the CUDA library is missing, importing fails loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libldn.so")


class LdnError(RuntimeError):
    pass


class ldn_tensor(C.Structure):
    _fields_ = [
        ("name", C.c_char_p),
        ("data", C.c_void_p),
        ("dtype", C.c_int),
        ("ndim", C.c_int),
        ("shape", C.c_int64 * 4),
    ]


class ldn_config(C.Structure):
    _fields_ = [
        ("max_rows", C.c_int),
        ("max_h", C.c_int),
        ("max_w", C.c_int),
        ("max_ctx_tokens", C.c_int),
        ("use_graph", C.c_int),
    ]


_p, _i, _l, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes; every entry returns int except the ones listed in _RESTYPES
SIGNATURES = {
    "ldn_last_error": [],
    "ldn_version": [],
    "ldn_create": [C.POINTER(ldn_config), C.POINTER(_p)],
    "ldn_destroy": [_p],
    "ldn_load_weights": [_p, _i, C.POINTER(ldn_tensor), _i, _p],
    "ldn_set_sigmas": [_p, C.POINTER(C.c_float), C.POINTER(C.c_float), _i],
    "ldn_set_context": [_p, _p, _i, _i, _p],
    "ldn_unet_denoise": [_p, _p, _p, _p, _i, _i, _i, _p],
    "ldn_unet_last_launches": [_p],
    "ldn_cfg_step": [_p, _p, _p, _f, _i, _f, _f, _f, _p, _p, _p, _l, _p],
    "ldn_resample_bilinear": [_p, _p, _i, _i, _i, _i, _i, _p],
    "ldn_bislerp": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "ldn_vae_decode": [_p, _p, _p, _i, _i, _i, _p],
    "ldn_vae_encode": [_p, _p, _p, _i, _i, _i, _p],
    "ldn_taesd_decode": [_p, _p, _p, _i, _i, _i, _p],
    "ldn_flux_forward": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "ldn_clip_encode": [_p, _p, _i, _p, _p, _p],
    "ldn_clip_set_extra_embeddings": [_p, _p, _i, _p],
    "ldn_t5_encode": [_p, _p, _p, _i, _i, _p, _p],
    "ldn_gemm_bf16": [_p, _l, _i, _p, _l, _i, _p, _i, _i, _p, _p, _i, _i, _p, _l, _p, _l, _p, _i, _i, _i, _i, _p],
    "ldn_conv3x3_bf16": [_p, _p, _i, _i, _i, _i, _i, _p, _p, _i, _p, _p, _p],
    "ldn_attention_bf16": [_p, _l, _p, _l, _p, _l, _l, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _p, _l, _p],
    "ldn_groupnorm_bf16": [_p, _i, _p, _i, _i, _i, _i, _f, _p, _p, _i, _p, _p],
    "ldn_layernorm_bf16": [_p, _i, _i, _f, _p, _p, _p, _p],
    "ldn_conv3x3_groupnorm_bf16": [_p, _p, _i, _i, _i, _i, _i, _p, _p, _i, _p, _f, _p, _p, _i, _p, _p, _p, _p],
}
_RESTYPES = {"ldn_last_error": C.c_char_p, "ldn_destroy": None}

_lib = None
MISSING: list[str] = []


def load() -> C.CDLL:
    """dlopen libldn.so and attach prototypes.  Raises if the library or any declared symbol is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LdnError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU/PyTorch fallback for the hot path)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            MISSING.append(name)  # tests assert this stays empty; calling a missing entry raises AttributeError
            continue
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().ldn_last_error()
        raise LdnError(msg.decode() if msg else f"libldn error {rc}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (or None -> NULL)."""
    return None if t is None else t.data_ptr()


def cur_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream
