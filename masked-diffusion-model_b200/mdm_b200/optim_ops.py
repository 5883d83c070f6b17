"""ctypes wrappers of the fused optimiser tail (csrc/optim.cu)."""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int, c_void_p

from . import _lib
from ._lib import check, lib, ptr, stream_ptr

_P = c_void_p
_lib.register({
    "mdm_grad_sumsq": (c_int, [_P, ctypes.c_int64, _P, _P, _P]),
    "mdm_adam_ema_step": (c_int, [_P, _P, _P, _P, _P, _P, ctypes.c_int64, _P] + [c_float] * 10 + [c_int, _P]),
    "mdm_adam_ema_step_dev": (c_int, [_P, _P, _P, _P, _P, _P, ctypes.c_int64, _P, _P] + [c_float] * 6 + [c_int, _P]),
})

MODE = {"adam": 0, "adamw": 1, "sgd": 2}


def grad_sumsq(g, ws, out):
    check(lib().mdm_grad_sumsq(ptr(g), g.numel(), ptr(ws), ptr(out), stream_ptr(g.device)))


def adam_ema_step(p, g, m, v, ema, p16, gnorm_sq, lr, beta1, beta2, eps, wd, bc1, bc2, max_norm, ema_decay, grad_scale, mode):
    check(lib().mdm_adam_ema_step(ptr(p), ptr(g), ptr(m), ptr(v), ptr(ema), ptr(p16), p.numel(), ptr(gnorm_sq),
                                  lr, beta1, beta2, eps, wd, bc1, bc2, max_norm, ema_decay, grad_scale, mode,
                                  stream_ptr(p.device)))


def adam_ema_step_dev(p, g, m, v, ema, p16, gnorm_sq, hyper, beta1, beta2, eps, wd, max_norm, grad_scale, mode):
    """hyper: device float32[4] = (lr, bias_c1, bias_c2, ema_decay) -- CUDA-graph replayable"""
    check(lib().mdm_adam_ema_step_dev(ptr(p), ptr(g), ptr(m), ptr(v), ptr(ema), ptr(p16), p.numel(), ptr(gnorm_sq),
                                      ptr(hyper), beta1, beta2, eps, wd, max_norm, grad_scale, mode,
                                      stream_ptr(p.device)))
