"""ctypes wrappers of the fused optimiser tail (csrc/optim.cu)."""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int, c_void_p

from . import _lib
from ._lib import check, lib, ptr, stream_ptr

_P = c_void_p
_D = ctypes.c_double
_lib.register({
    "mdm_grad_sumsq": (c_int, [_P, ctypes.c_int64, _P, _P, _P]),
    "mdm_publish_stats": (c_int, [_P, c_int, _P, _P, _P]),
    "mdm_adam_ema_step": (c_int, [_P, _P, _P, _P, _P, _P, ctypes.c_int64, _P, c_float, _D, _D, c_float, c_float, _D, _D,
                                  c_float, c_float, c_float, c_int, _P]),
    "mdm_adam_ema_step_dev": (c_int, [_P, _P, _P, _P, _P, _P, ctypes.c_int64, _P, _P, _D, _D, c_float, c_float, c_float, c_float,
                                      c_int, _P]),
})

MODE = {"adam": 0, "adamw": 1, "sgd": 2}


def grad_sumsq(g, ws, out):
    check(lib().mdm_grad_sumsq(ptr(g), g.numel(), ptr(ws), ptr(out), stream_ptr(g.device)))


def adam_ema_step(p, g, m, v, ema, p16, gnorm_sq, lr, beta1, beta2, eps, wd, bc1, bc2, max_norm, ema_decay, grad_scale, mode):
    check(lib().mdm_adam_ema_step(ptr(p), ptr(g), ptr(m), ptr(v), ptr(ema), ptr(p16), p.numel(), ptr(gnorm_sq),
                                  lr, beta1, beta2, eps, wd, bc1, bc2, max_norm, ema_decay, grad_scale, mode,
                                  stream_ptr(p.device)))


def adam_ema_step_dev(p, g, m, v, ema, p16, gnorm_sq, hyper, beta1, beta2, eps, wd, max_norm, grad_scale, mode):
    """hyper: device float32[4] = (lr, lr / bias_c1, sqrt(bias_c2), ema_decay) -- CUDA-graph replayable"""
    check(lib().mdm_adam_ema_step_dev(ptr(p), ptr(g), ptr(m), ptr(v), ptr(ema), ptr(p16), p.numel(), ptr(gnorm_sq),
                                      ptr(hyper), beta1, beta2, eps, wd, max_norm, grad_scale, mode,
                                      stream_ptr(p.device)))


def publish_stats(src, counter, dst_pinned):
    """src: device float32[n]; counter: device int32[1]; dst_pinned: pinned host float32[>= n + 1] (device-mapped)"""
    check(lib().mdm_publish_stats(ptr(src), src.numel(), ptr(counter), ctypes.c_void_p(dst_pinned.data_ptr()),
                                  stream_ptr(src.device)))
