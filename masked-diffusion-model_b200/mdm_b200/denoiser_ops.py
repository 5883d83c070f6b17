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
        ("bias", c_void_p), ("bias2", c_void_p), ("rowvec", c_void_p), ("ld_rowvec", c_longlong),
        ("resid", c_void_p), ("ld_resid", c_longlong), ("accumulate", c_int),
        ("y_f32", c_void_p),
        ("x2", c_void_p), ("ld_x2", c_longlong), ("cin2", c_int), ("w2", c_void_p),
        ("dw", c_void_p), ("w_col0", c_longlong), ("w_cols", c_int),
        ("dbias", c_void_p), ("dbias2", c_void_p),
        ("splitk_ws", c_void_p), ("splitk_ws_floats", c_longlong),
        ("qsum", c_void_p),
        ("up2x", c_int), ("up_a", c_int), ("up_b", c_int),
        ("gn_coef", c_void_p),
    ]


_P = c_void_p
_LL = c_longlong
_I64 = ctypes.c_int64
_lib.register({
    "mdm_conv_fprop": (c_int, [POINTER(ConvArgs), _P]),
    "mdm_conv_dgrad": (c_int, [POINTER(ConvArgs), _P]),
    "mdm_conv_wgrad": (c_int, [POINTER(ConvArgs), _P]),
    "mdm_up2x_weights": (c_int, [_P, _P, c_int, c_int, _P]),
    "mdm_gn_coef_q": (c_int, [_P, c_int, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, _P]),
    "mdm_reserve_sms": (c_int, [c_int]),
    "mdm_set_sched_workspace": (c_int, [_P, c_int]),
    "mdm_gn_ws_floats": (_I64, [c_int, c_int, c_int, c_int]),
    "mdm_gn_fwd_kind": (c_int, [c_int, c_int, c_int, c_int]),
    "mdm_gn_silu_fwd_q": (c_int, [_P, _LL, _P, _LL, _P, _P, _P, _P, c_int, _P, c_int, c_int, c_int, c_int, c_float, c_int, _P]),
    "mdm_gn_silu_fwd": (c_int, [_P, _LL, _P, _LL, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, c_int, _P]),
    "mdm_gn_silu_bwd": (c_int, [_P, _LL, _P, _LL, _P, _LL, _P, _LL, _P, _LL, _P, _P, _P, _P, _P, _P, _P, _LL, _P,
                                c_int, c_int, c_int, c_int, c_int, _P]),
    "mdm_conv_in_fwd": (c_int, [_P, _P, _P, _P, _LL, c_int, c_int, c_int, c_int, c_int, _P]),
    "mdm_conv_in_wgrad": (c_int, [_P, _P, _LL, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "mdm_conv_out_fwd": (c_int, [_P, _LL, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "mdm_conv_out_bwd": (c_int, [_P, _LL, _P, _P, _P, _LL, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "mdm_im2col3x3": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "mdm_tapsum3x3": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "mdm_upsample2x_fwd": (c_int, [_P, _LL, _P, _LL, c_int, c_int, c_int, c_int, _P]),
    "mdm_upsample2x_bwd": (c_int, [_P, _LL, _P, _LL, c_int, c_int, c_int, c_int, _P]),
    "mdm_zero_insert2x": (c_int, [_P, _LL, _P, _LL, c_int, c_int, c_int, c_int, _P]),
    "mdm_attention_fwd": (c_int, [_P, _P, c_int, c_int, c_int, _P]),
    "mdm_attention_bwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P]),
    "mdm_timestep_embedding": (c_int, [_P, _P, c_int, c_int, _P]),
    "mdm_silu_fwd": (c_int, [_P, _P, _I64, _P]),
    "mdm_silu_bwd": (c_int, [_P, _P, _P, _I64, _P]),
    "mdm_colsum": (c_int, [_P, _LL, _P, _P, _I64, c_int, _P]),
    "mdm_sample_colsum": (c_int, [_P, _LL, _P, _LL, _P, c_int, c_int, c_int, _P]),
    "mdm_mse_residual": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _P]),
    "mdm_cast_f32_bf16": (c_int, [_P, _P, _I64, _P]),
})


def _dp(t):
    return t.data_ptr() if t is not None else None


# zeroed fp32 scratch per device: small-M layers split their K loop over the SMs through it (the kernels
# leave it zeroed again, and launches are stream-ordered, so one buffer serves every layer)
SPLITK_WS_FLOATS = 1 << 20
_splitk_ws = {}


_sched_ws = {}


def _ensure_sched_ws(device):
    """zeroed counters for the GEMM kernel's optional dynamic work distribution (MDM_IGEMM_DYNAMIC=1), lent once per
    process (one process per GPU)"""
    if not _sched_ws:
        ws = _sched_ws[device] = torch.zeros(4096, dtype=torch.int32, device=device)
        check(lib().mdm_set_sched_workspace(ws.data_ptr(), ws.numel()))


def _splitk(a, device):
    _ensure_sched_ws(device)
    # one workspace per (device, stream): GEMMs on different streams may run concurrently
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _splitk_ws.get(key)
    if ws is None:
        ws = _splitk_ws[key] = torch.zeros(SPLITK_WS_FLOATS, dtype=torch.float32, device=device)
    a.splitk_ws, a.splitk_ws_floats = ws.data_ptr(), ws.numel()


def pix_ld(t: torch.Tensor) -> int:
    """channel stride (elements between consecutive pixels) of an NHWC / [rows, C] view"""
    assert t.stride(-1) == 1 or t.shape[-1] == 1
    for d in range(t.dim() - 2, -1, -1):
        if t.shape[d] > 1:
            return t.stride(d)
    return t.shape[-1]


def conv_fprop(x, w, y, N, H, W, ksize=3, stride=1, bias=None, rowvec=None, resid=None, accumulate=False,
               y_f32=None, x2=None, w2=None, bias2=None, cout=None, ld_rowvec=None, qsum=None, gn_coef=None):
    """y[N,H,W,cout] = conv(x, w) (+bias +rowvec[n] +resid) ; x: [N,H*s,W*s,cin] view, w packed
    bf16 [cout, k*k, cin]; optional fused 1x1 shortcut (x2, w2)."""
    a = ConvArgs()
    a.x, a.ld_x, a.cin = _dp(x), pix_ld(x), x.shape[-1]
    a.w = _dp(w)
    if y is not None:
        a.y, a.ld_y, a.cout = _dp(y), pix_ld(y), y.shape[-1]
    else:
        a.y, a.ld_y, a.cout = None, 8, cout
    a.N, a.H, a.W, a.ksize, a.stride = N, H, W, ksize, stride
    a.bias = _dp(bias)
    a.bias2 = _dp(bias2)
    if rowvec is not None:
        a.rowvec, a.ld_rowvec = _dp(rowvec), (ld_rowvec if ld_rowvec is not None else rowvec.stride(0))
    if resid is not None:
        a.resid, a.ld_resid = _dp(resid), pix_ld(resid)
    a.accumulate = int(accumulate)
    a.y_f32 = _dp(y_f32)
    if x2 is not None:
        a.x2, a.ld_x2, a.cin2, a.w2 = _dp(x2), pix_ld(x2), x2.shape[-1], _dp(w2)
    a.qsum = _dp(qsum)          # fused GroupNorm statistics of the output: qsum[N, cout/4, 2] += quad (sum, sumsq)
    a.gn_coef = _dp(gn_coef)    # x is the RAW activation: GroupNorm + SiLU applied while the halo tiles are filled (inference)
    _splitk(a, x.device)
    check(lib().mdm_conv_fprop(ctypes.byref(a), stream_ptr(x.device)))


def gn_coef_q(qa, qb, gamma, beta, coef, N, HW, C, G, eps, stats=None):
    """coef[N, C, 2] = (scale, shift) of a GroupNorm site from its producers' quad sums (qa [N, Ca/4, 2], qb for the second
    half of a concatenation or None)"""
    check(lib().mdm_gn_coef_q(ptr(qa), qa.shape[1], ptr(qb), ptr(gamma), ptr(beta), ptr(coef), ptr(stats), N, HW, C, G, eps,
                              stream_ptr(coef.device)))


def up2x_weights(w32, out_bf16, cout, cin):
    """parity weights of the fused upsample convolution: w32 [cout, 9, cin] fp32 -> out [4, cout, 4, cin] bf16"""
    check(lib().mdm_up2x_weights(ptr(w32), ptr(out_bf16), cout, cin, stream_ptr(w32.device)))


def conv_up2x_fprop(x, w4, y, N, H, W, bias=None, qsum=None):
    """y[N, 2H, 2W, cout] = conv3x3(nearest_upsample_2x(x[N, H, W, cin])) + bias without building the upsampled image:
    four launches, one per output parity, each a 2x2 convolution of x with the pre-summed weights w4[2a+b] (up2x_weights)"""
    for pa in range(2):
        for pb in range(2):
            a = ConvArgs()
            a.x, a.ld_x, a.cin = _dp(x), pix_ld(x), x.shape[-1]
            a.w = _dp(w4[2 * pa + pb])
            a.y, a.ld_y, a.cout = _dp(y), pix_ld(y), y.shape[-1]
            a.N, a.H, a.W, a.ksize, a.stride = N, H, W, 3, 1
            a.bias = _dp(bias)
            a.qsum = _dp(qsum)
            a.up2x, a.up_a, a.up_b = 1, pa, pb
            _ensure_sched_ws(x.device)
            check(lib().mdm_conv_fprop(ctypes.byref(a), stream_ptr(x.device)))


def reserve_sms(n):
    """leave n SMs out of the persistent GEMM grids (a concurrent kernel on another stream owns them); returns the
    previous reservation (n < 0: query only)"""
    return int(lib().mdm_reserve_sms(int(n)))


def conv_dgrad(dy, w, dx, N, H, W, ksize=3, resid=None, accumulate=False, dx_f32=None, cin=None):
    """dx[N,H,W,cin] (+)= dgrad(dy[N,H,W,cout], w[cout,k*k,cin]) for a stride-1 layer."""
    a = ConvArgs()
    a.x, a.ld_x, a.cout = _dp(dy), pix_ld(dy), dy.shape[-1]
    a.w = _dp(w)
    if dx is not None:
        a.y, a.ld_y, a.cin = _dp(dx), pix_ld(dx), dx.shape[-1]
    else:
        a.y, a.ld_y, a.cin = None, 8, cin
    a.N, a.H, a.W, a.ksize, a.stride = N, H, W, ksize, 1
    if resid is not None:
        a.resid, a.ld_resid = _dp(resid), pix_ld(resid)
    a.accumulate = int(accumulate)
    a.y_f32 = _dp(dx_f32)
    a.w_cols = w.shape[-1]
    _splitk(a, dy.device)
    check(lib().mdm_conv_dgrad(ctypes.byref(a), stream_ptr(dy.device)))


def conv_wgrad(x, dy, dw, N, H, W, ksize=3, stride=1, dbias=None, dbias2=None):
    """dw[cout,k*k,cin] (fp32) += wgrad(x[N,H*s,W*s,cin], dy[N,H,W,cout]); optional dbias[cout] (+dbias2) += column
    sums of dy (bias gradient), fused into the same kernel"""
    a = ConvArgs()
    a.x, a.ld_x, a.cin = _dp(x), pix_ld(x), x.shape[-1]
    a.y, a.ld_y, a.cout = _dp(dy), pix_ld(dy), dy.shape[-1]
    a.N, a.H, a.W, a.ksize, a.stride = N, H, W, ksize, stride
    a.dw = _dp(dw)
    a.dbias, a.dbias2 = _dp(dbias), _dp(dbias2)
    a.w_cols = dw.shape[-1]
    _ensure_sched_ws(dw.device)
    check(lib().mdm_conv_wgrad(ctypes.byref(a), stream_ptr(dw.device)))


def pack_conv_weight(w_nchw: torch.Tensor) -> torch.Tensor:
    """diffusers (cout, cin, kh, kw) fp32 -> packed (cout, kh*kw, cin) (same dtype)"""
    co, ci, kh, kw = w_nchw.shape
    return w_nchw.permute(0, 2, 3, 1).reshape(co, kh * kw, ci).contiguous()


def unpack_conv_weight(w_packed: torch.Tensor, k: int) -> torch.Tensor:
    co, taps, ci = w_packed.shape
    return w_packed.reshape(co, k, k, ci).permute(0, 3, 1, 2).contiguous()


# ---- non-GEMM kernels -------------------------------------------------------------------------------
def _s(t):
    return stream_ptr(t.device)


def gn_ws_floats(N, HW, C, G=32):
    return int(lib().mdm_gn_ws_floats(N, HW, C, G))


def gn_silu_fwd(x, y, gamma, beta, stats, ws, N, HW, C, G=32, eps=1e-5, silu=True):
    check(lib().mdm_gn_silu_fwd(_dp(x), pix_ld(x), _dp(y), pix_ld(y), _dp(gamma), _dp(beta), _dp(stats), _dp(ws),
                                N, HW, C, G, eps, int(silu), _s(x)))


def gn_fwd_kind(N, HW, C, G=32):
    """0: one CTA per sample, 1: one cluster per sample, 2: statistics pass + apply pass (fused statistics save a pass)"""
    return lib().mdm_gn_fwd_kind(N, HW, C, G)


def gn_silu_fwd_q(x, y, gamma, beta, stats, qa, qb, N, HW, C, G=32, eps=1e-5, silu=True):
    """GroupNorm(+SiLU) forward from quad sums produced by the convolutions that wrote x (qa: the first qa.shape[1]
    quads, qb: the rest, None when qa covers all channels): x is read once"""
    check(lib().mdm_gn_silu_fwd_q(_dp(x), pix_ld(x), _dp(y), pix_ld(y), _dp(gamma), _dp(beta), _dp(stats), _dp(qa),
                                  qa.shape[1], _dp(qb), N, HW, C, G, eps, int(silu), _s(x)))


def gn_silu_bwd(x, dy, dx, gamma, beta, stats, dgamma, dbeta, ws, N, HW, C, G=32, silu=True, add=None, add2=None,
                colsum=None, ld_colsum=0, dbias=None):
    """dx = GroupNorm(+SiLU) backward (+ add + add2); optional fused per-sample column sums of the GroupNorm
    part of dx: colsum[n, c] += ..., dbias[c] += ... (fp32, accumulated)"""
    check(lib().mdm_gn_silu_bwd(_dp(x), pix_ld(x), _dp(dy), pix_ld(dy),
                                _dp(add), pix_ld(add) if add is not None else 0,
                                _dp(add2), pix_ld(add2) if add2 is not None else 0,
                                _dp(dx), pix_ld(dx), _dp(gamma), _dp(beta), _dp(stats), _dp(dgamma), _dp(dbeta),
                                _dp(ws), _dp(colsum), ld_colsum, _dp(dbias), N, HW, C, G, int(silu), _s(x)))


def conv_in_fwd(img, w, bias, y, N, C, H, W, cout):
    check(lib().mdm_conv_in_fwd(_dp(img), _dp(w), _dp(bias), _dp(y), pix_ld(y), N, C, H, W, cout, _s(img)))


def conv_in_wgrad(img, dy, dw, dbias, N, C, H, W, cout):
    check(lib().mdm_conv_in_wgrad(_dp(img), _dp(dy), pix_ld(dy), _dp(dw), _dp(dbias), N, C, H, W, cout, _s(img)))


def conv_out_fwd(x, w, bias, y, N, C, H, W, cin):
    check(lib().mdm_conv_out_fwd(_dp(x), pix_ld(x), _dp(w), _dp(bias), _dp(y), N, C, H, W, cin, _s(x)))


def conv_out_bwd(x, w, dy, dx, dw, dbias, N, C, H, W, cin):
    check(lib().mdm_conv_out_bwd(_dp(x), pix_ld(x), _dp(w), _dp(dy), _dp(dx), pix_ld(dx) if dx is not None else 0,
                                 _dp(dw), _dp(dbias), N, C, H, W, cin, _s(x)))


def im2col3x3(img, out, N, C, H, W, flip=False):
    """planar fp32 [N,C,H,W] -> bf16 [N*H*W, 64], column tap*C + c (zero padded); flip mirrors the taps"""
    check(lib().mdm_im2col3x3(_dp(img), _dp(out), N, C, H, W, int(flip), _s(img)))


def tapsum3x3(z, bias, out, N, C, H, W):
    """out[n,c,h,w] = bias[c] + sum_tap z[(n,h+dh,w+dw), tap*C + c]; z fp32 [N*H*W, 32]"""
    check(lib().mdm_tapsum3x3(_dp(z), _dp(bias), _dp(out), N, C, H, W, _s(z)))


def upsample2x_fwd(x, y, N, H, W, C):
    check(lib().mdm_upsample2x_fwd(_dp(x), pix_ld(x), _dp(y), pix_ld(y), N, H, W, C, _s(x)))


def upsample2x_bwd(dy, dx, N, H, W, C):
    check(lib().mdm_upsample2x_bwd(_dp(dy), pix_ld(dy), _dp(dx), pix_ld(dx), N, H, W, C, _s(dy)))


def zero_insert2x(x, y, N, H, W, C):
    check(lib().mdm_zero_insert2x(_dp(x), pix_ld(x), _dp(y), pix_ld(y), N, H, W, C, _s(x)))


def attention_fwd(qkv, out, N, L, C):
    check(lib().mdm_attention_fwd(_dp(qkv), _dp(out), N, L, C, _s(qkv)))


def attention_bwd(qkv, dout, dqkv, N, L, C):
    check(lib().mdm_attention_bwd(_dp(qkv), _dp(dout), _dp(dqkv), N, L, C, _s(qkv)))


def timestep_embedding(t, out, N, dim):
    check(lib().mdm_timestep_embedding(_dp(t), _dp(out), N, dim, _s(t)))


def silu_fwd(x_f32, y_bf16):
    check(lib().mdm_silu_fwd(_dp(x_f32), _dp(y_bf16), x_f32.numel(), _s(x_f32)))


def silu_bwd(x_f32, dy_f32, dx_bf16):
    check(lib().mdm_silu_bwd(_dp(x_f32), _dp(dy_f32), _dp(dx_bf16), x_f32.numel(), _s(x_f32)))


def colsum(dy, out, rows, C, out2=None):
    check(lib().mdm_colsum(_dp(dy), pix_ld(dy), _dp(out), _dp(out2), rows, C, _s(dy)))


def sample_colsum(dy, out, ld_out, dbias, N, HW, C):
    check(lib().mdm_sample_colsum(_dp(dy), pix_ld(dy), _dp(out), ld_out, _dp(dbias), N, HW, C, _s(dy)))


def mse_residual(x_in, net, shift, x0, weight, dnet, recon, loss, ws, per_sample):
    check(lib().mdm_mse_residual(_dp(x_in), _dp(net), _dp(shift), _dp(x0), _dp(weight), _dp(dnet), _dp(recon), _dp(loss),
                                 _dp(ws), per_sample, x_in.numel(), _s(x_in)))


def cast_f32_bf16(x, y):
    check(lib().mdm_cast_f32_bf16(_dp(x), _dp(y), x.numel(), _s(x)))
