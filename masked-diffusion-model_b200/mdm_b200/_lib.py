"""ctypes binding of libmdm_sm100.so (C ABI declared in include/mdm.h).

There is no CPU fallback: if the library is missing, `lib()` raises, and compute wrappers raise
when handed a non-CUDA tensor."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_uint32, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MDM_LIB_PATH") or os.path.join(_HERE, "libmdm_sm100.so")   # override: A/B builds of the kernels
HEADER_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "include", "mdm.h"))

MDM_F32, MDM_BF16, MDM_U8 = 0, 1, 2
FILL_CONST, FILL_DEGRADED_AREA, FILL_NON_DEGRADED = 0, 1, 2
AREA_IMAGE, AREA_CHANNEL = 0, 1
RNG_WORDS = 625
RNG_PAR_WORDS = 1280          # state buffers are allocated this long: words 640.. stage the advanced state

_P = c_void_p
_SIGS = {
    "mdm_last_error": (c_char_p, []),
    "mdm_version": (c_int, []),
    "mdm_device_available": (c_int, []),
    "mdm_launch_count": (ctypes.c_longlong, []),
    "mdm_rng_seed_host": (c_int, [_P, c_uint32]),
    "mdm_rng_jump_table_host": (c_int, [_P, c_int, c_int]),
    "mdm_rng_enable_parallel": (c_int, [_P, c_int, c_int]),
    "mdm_rng_set_par_stride": (c_int, [c_int]),
    "mdm_rng_advance_host": (c_int, [_P, c_int64, _P]),
    "mdm_rng_raw": (c_int, [_P, _P, c_int64, _P]),
    "mdm_rng_skip": (c_int, [_P, c_int64, _P]),
    "mdm_rng_uniform": (c_int, [_P, _P, c_int64, c_float, c_float, _P]),
    "mdm_rng_normal": (c_int, [_P, _P, c_int, c_int64, c_float, c_float, _P, _P]),
    "mdm_rng_threshold_mask": (c_int, [_P, _P, _P, _P, _P, c_int, c_int64, _P]),
    "mdm_rng_randperm_mask": (c_int, [_P, _P, _P, _P, c_int, c_int, _P]),
    "mdm_rng_randint": (c_int, [_P, _P, c_int64, c_int64, c_int64, _P]),
    "mdm_degrade_ws_floats": (c_int64, [c_int, c_int, c_int]),
    "mdm_degrade": (c_int, [_P, c_int, _P, c_int, c_int, c_float, c_int, _P, _P, _P, _P, _P,
                            c_int, c_int, c_int, _P]),
    "mdm_degrade_u8": (c_int, [_P, _P, c_int, c_int, c_float, c_int, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "mdm_sampler_step": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, _P, _P, c_int, c_int,
                                 c_float, c_int, c_int, c_int, _P, c_int64, c_int64, c_int64,
                                 _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "mdm_image_grid": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, _P, _P, _P]),
    "mdm_add_shift": (c_int, [_P, _P, c_int64, c_int64, c_int64, _P, c_int, c_int, c_int, _P]),
}

_lib = None


def register(sigs: dict):
    """Other modules (denoiser, optimiser) add their entry points here before first use."""
    _SIGS.update(sigs)
    if _lib is not None:
        _declare(_lib, sigs)


def _declare(l, sigs):
    for name, (res, args) in sigs.items():
        fn = getattr(l, name)
        fn.restype = res
        fn.argtypes = args


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        l = ctypes.CDLL(LIB_PATH)
        _declare(l, _SIGS)
        _lib = l
    return _lib


def check(rc: int):
    if rc != 0:
        raise RuntimeError(f"libmdm_sm100 error {rc}: {lib().mdm_last_error().decode()}")


def ptr(t):
    if t is None:
        return None
    return c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t: torch.Tensor, what="tensor"):
    if not t.is_cuda:
        raise RuntimeError(f"{what} must live on a CUDA device: this is the B200 path, there is no CPU fallback")
    return t
