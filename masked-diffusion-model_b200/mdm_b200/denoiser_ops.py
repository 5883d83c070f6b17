"""ctypes wrappers of the denoiser kernels (include/mdm.h, csrc/igemm.cu, csrc/nn_kernels.cu).

Tensors are NHWC bf16 "views": (tensor, ld) where `tensor` is a [N, H, W, C] (or [rows, C])
torch view whose last-dim stride is 1 and whose pixel stride is `ld`."""
from __future__ import annotations

import ctypes
from ctypes import POINTER, Structure, c_float, c_int, c_longlong, c_void_p

import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


class ConvArgs(Structure):
    _fields_ = [
        ("x", c_void_p), ("ld_x", c_longlong), ("cin", c_int),
        ("w", c_void_p),
        ("y", c_void_p), ("ld_y", c_longlong), ("cout", c_int),
        ("N", c_int), ("H", c_int), ("W", c_int),
        ("ksize", c_int), ("stride", c_int),
        ("bias", c_void_p), ("rowvec", c_void_p), ("ld_rowvec", c_longlong),
        ("resid", c_void_p), ("ld_resid", c_longlong), ("accumulate", c_int),
        ("y_f32", c_void_p),
        ("x2", c_void_p), ("ld_x2", c_longlong), ("cin2", c_int), ("w2", c_void_p),
        ("dw", c_void_p), ("w_col0", c_longlong), ("w_cols", c_int),
    ]


_P = c_void_p
_lib.register({
    "mdm_conv_fprop": (c_int, [POINTER(ConvArgs), _P]),
    "mdm_conv_dgrad": (c_int, [POINTER(ConvArgs), _P]),
    "mdm_conv_wgrad": (c_int, [POINTER(ConvArgs), _P]),
})


def _dp(t):
    return t.data_ptr() if t is not None else None


def pix_ld(t: torch.Tensor) -> int:
    """channel stride (elements between consecutive pixels) of an NHWC / [rows, C] view"""
    assert t.stride(-1) == 1 or t.shape[-1] == 1
    for d in range(t.dim() - 2, -1, -1):
        if t.shape[d] > 1:
            return t.stride(d)
    return t.shape[-1]


def conv_fprop(x, w, y, N, H, W, ksize=3, stride=1, bias=None, rowvec=None, resid=None, accumulate=False,
               y_f32=None, x2=None, w2=None):
    """y[N,H,W,cout] = conv(x, w) (+bias +rowvec[n] +resid) ; x: [N,H*s,W*s,cin] view, w packed
    bf16 [cout, k*k, cin]; optional fused 1x1 shortcut (x2, w2)."""
    a = ConvArgs()
    a.x, a.ld_x, a.cin = _dp(x), pix_ld(x), x.shape[-1]
    a.w = _dp(w)
    a.y, a.ld_y, a.cout = _dp(y), pix_ld(y), y.shape[-1]
    a.N, a.H, a.W, a.ksize, a.stride = N, H, W, ksize, stride
    a.bias = _dp(bias)
    if rowvec is not None:
        a.rowvec, a.ld_rowvec = _dp(rowvec), rowvec.stride(0)
    if resid is not None:
        a.resid, a.ld_resid = _dp(resid), pix_ld(resid)
    a.accumulate = int(accumulate)
    a.y_f32 = _dp(y_f32)
    if x2 is not None:
        a.x2, a.ld_x2, a.cin2, a.w2 = _dp(x2), pix_ld(x2), x2.shape[-1], _dp(w2)
    check(lib().mdm_conv_fprop(ctypes.byref(a), stream_ptr(y.device)))


def conv_dgrad(dy, w, dx, N, H, W, ksize=3, resid=None, accumulate=False, dx_f32=None):
    """dx[N,H,W,cin] (+)= dgrad(dy[N,H,W,cout], w[cout,k*k,cin]) for a stride-1 layer."""
    a = ConvArgs()
    a.x, a.ld_x, a.cout = _dp(dy), pix_ld(dy), dy.shape[-1]
    a.w = _dp(w)
    a.y, a.ld_y, a.cin = _dp(dx), pix_ld(dx), dx.shape[-1]
    a.N, a.H, a.W, a.ksize, a.stride = N, H, W, ksize, 1
    if resid is not None:
        a.resid, a.ld_resid = _dp(resid), pix_ld(resid)
    a.accumulate = int(accumulate)
    a.y_f32 = _dp(dx_f32)
    a.w_cols = w.shape[-1]
    check(lib().mdm_conv_dgrad(ctypes.byref(a), stream_ptr(dx.device)))


def conv_wgrad(x, dy, dw, N, H, W, ksize=3, stride=1):
    """dw[cout,k*k,cin] (fp32) += wgrad(x[N,H*s,W*s,cin], dy[N,H,W,cout])"""
    a = ConvArgs()
    a.x, a.ld_x, a.cin = _dp(x), pix_ld(x), x.shape[-1]
    a.y, a.ld_y, a.cout = _dp(dy), pix_ld(dy), dy.shape[-1]
    a.N, a.H, a.W, a.ksize, a.stride = N, H, W, ksize, stride
    a.dw = _dp(dw)
    a.w_cols = dw.shape[-1]
    check(lib().mdm_conv_wgrad(ctypes.byref(a), stream_ptr(dw.device)))


def pack_conv_weight(w_nchw: torch.Tensor) -> torch.Tensor:
    """diffusers (cout, cin, kh, kw) fp32 -> packed (cout, kh*kw, cin) (same dtype)"""
    co, ci, kh, kw = w_nchw.shape
    return w_nchw.permute(0, 2, 3, 1).reshape(co, kh * kw, ci).contiguous()


def unpack_conv_weight(w_packed: torch.Tensor, k: int) -> torch.Tensor:
    co, taps, ci = w_packed.shape
    return w_packed.reshape(co, k, k, ci).permute(0, 3, 1, 2).contiguous()
