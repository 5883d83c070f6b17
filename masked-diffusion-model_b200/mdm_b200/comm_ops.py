"""ctypes wrappers of the peer-memory gradient all-reduce (csrc/allreduce.cu) and its CUDA-IPC plumbing."""
from __future__ import annotations

import ctypes
from ctypes import Structure, c_int, c_int64, c_void_p

import torch

from . import _lib
from ._lib import check, lib, stream_ptr

MAX_RANKS = 8


class P2PCommStruct(Structure):
    _fields_ = [("buf", c_void_p * MAX_RANKS), ("flag", c_void_p * MAX_RANKS), ("rank", c_int), ("world", c_int)]


_P = c_void_p
_lib.register({
    "mdm_p2p_flag_words": (c_int, []),
    "mdm_ipc_export": (c_int, [_P, _P, ctypes.POINTER(c_int64)]),
    "mdm_ipc_open": (c_int, [_P, c_int64, ctypes.POINTER(c_void_p)]),
    "mdm_ipc_close": (c_int, [_P, c_int64]),
    "mdm_p2p_allreduce": (c_int, [ctypes.POINTER(P2PCommStruct), c_int64, c_int64, c_int, _P]),
    "mdm_p2p_barrier": (c_int, [ctypes.POINTER(P2PCommStruct), c_int, _P]),
    "mdm_memcpy_async": (c_int, [_P, _P, c_int64, _P]),
    "mdm_reduce_slices": (c_int, [_P, _P, c_int64, c_int64, c_int, c_int, _P]),
})


def flag_words() -> int:
    return int(lib().mdm_p2p_flag_words())


def ipc_export(t: torch.Tensor):
    """-> (64-byte handle, byte offset of t inside its cudaMalloc allocation)"""
    h = ctypes.create_string_buffer(64)
    off = c_int64(0)
    check(lib().mdm_ipc_export(c_void_p(t.data_ptr()), h, ctypes.byref(off)))
    return bytes(h.raw), int(off.value)


def ipc_open(handle: bytes, offset: int) -> int:
    out = c_void_p(0)
    check(lib().mdm_ipc_open(ctypes.create_string_buffer(handle, 64), offset, ctypes.byref(out)))
    return int(out.value)


def p2p_allreduce(comm: P2PCommStruct, offset: int, count: int, blocks: int, device):
    check(lib().mdm_p2p_allreduce(ctypes.byref(comm), offset, count, blocks, stream_ptr(device)))


def p2p_barrier(comm: P2PCommStruct, which: int, device):
    check(lib().mdm_p2p_barrier(ctypes.byref(comm), which, stream_ptr(device)))


def memcpy_async(dst_ptr: int, src_ptr: int, nbytes: int, device):
    check(lib().mdm_memcpy_async(c_void_p(dst_ptr), c_void_p(src_ptr), nbytes, stream_ptr(device)))


def reduce_slices(slice_ptr: int, staging_ptr: int, stride: int, count: int, rank: int, world: int, device):
    check(lib().mdm_reduce_slices(c_void_p(slice_ptr), c_void_p(staging_ptr), stride, count, rank, world, stream_ptr(device)))


def ipc_close(ptr: int, offset: int):
    check(lib().mdm_ipc_close(c_void_p(ptr), offset))
