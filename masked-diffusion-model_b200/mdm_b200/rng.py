"""Device-resident shadow of torch's CPU mt19937 generator.

The reference draws every mask and every shift tensor from the CPU generator even when the
model is on the GPU (scheduler.py:282,288,434,440,620,675,707).  To give
bit-identical masks under the same seed, the generator state is copied to the device once
(`adopt_torch`), consumed there by the kernels in csrc/rng.cu, and written back to torch on
request (`release_to_torch`) so later CPU draws continue from the right position."""
from __future__ import annotations

import struct

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr

MT_N = 624


def parse_torch_state(state: torch.Tensor):
    """-> (seed, uint32[624], pos).  Layout: SURVEY.md section 3.2.1 (iii)."""
    raw = state.numpy().tobytes()
    seed, left, seeded, nxt = struct.unpack_from("<QiiQ", raw, 0)
    key = np.frombuffer(raw, dtype="<u8", count=MT_N, offset=24).astype(np.uint32)
    pos = MT_N if left == 1 else int(nxt)
    return int(seed), key, pos


def pack_torch_state(seed: int, key: np.ndarray, pos: int, template: torch.Tensor | None = None) -> torch.Tensor:
    buf = bytearray(template.numpy().tobytes()) if template is not None else bytearray(5056)
    struct.pack_into("<QiiQ", buf, 0, seed, 625 - pos, 1, pos)
    buf[24:24 + MT_N * 8] = key.astype("<u8").tobytes()
    return torch.frombuffer(buf, dtype=torch.uint8).clone()


# parallel stream generation (csrc/mt_jump.cu): one jump-polynomial table per process and device
JUMP_POLYS = 512            # CTAs per launch beyond the first: 513 x 64 blocks x 624 words = 20.5 M words per launch
JUMP_BLOCKS_PER_CTA = 64    # 39936 words per CTA (~30 us of generation behind a ~0.1 ms jump)
_jump_tables = {}


def ensure_parallel(device):
    """compute the jump polynomials on the host (Berlekamp-Massey + 512 modular products, ~0.5 s, once per process),
    copy them to `device` and lend them to the library; MDM_RNG_PARALLEL=0 keeps every draw on one CTA"""
    import os
    if os.environ.get("MDM_RNG_PARALLEL", "1") == "0":
        return False
    key = (device.type, device.index)
    if key not in _jump_tables:
        host = np.empty(JUMP_POLYS * MT_N, dtype=np.uint32)
        check(lib().mdm_rng_jump_table_host(host.ctypes.data, JUMP_POLYS, JUMP_BLOCKS_PER_CTA))
        t = torch.from_numpy(host.view(np.int32)).to(device)
        with torch.cuda.device(device):
            check(lib().mdm_rng_enable_parallel(t.data_ptr(), JUMP_POLYS, JUMP_BLOCKS_PER_CTA))
        _jump_tables.clear()            # the library holds ONE table (one process per GPU)
        _jump_tables[key] = t
    return True


class DeviceMT19937:
    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceMT19937 needs a CUDA device (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        # 1280 words: [0..624] key + position, [640..1264] staging of the advanced state (multi-CTA draws)
        self.state = torch.zeros(_lib.RNG_PAR_WORDS, dtype=torch.int32, device=self.device)
        ensure_parallel(self.device)
        self.seed = 0
        self._template = None
        self.words_drawn = 0

    # -- state movement ------------------------------------------------------------------------
    def _upload(self, key: np.ndarray, pos: int):
        host = np.empty(_lib.RNG_WORDS, dtype=np.uint32)
        host[:MT_N] = key
        host[MT_N] = pos
        self.state[:_lib.RNG_WORDS].copy_(torch.from_numpy(host.view(np.int32)))

    def manual_seed(self, seed: int):
        host = np.empty(_lib.RNG_WORDS, dtype=np.uint32)
        check(lib().mdm_rng_seed_host(host.ctypes.data, int(seed) & 0xFFFFFFFF))
        self.seed = int(seed)
        self.state[:_lib.RNG_WORDS].copy_(torch.from_numpy(host.view(np.int32)))
        return self

    def adopt_torch(self, generator: torch.Generator | None = None):
        st = generator.get_state() if generator is not None else torch.get_rng_state()
        self.seed, key, pos = parse_torch_state(st)
        self._template = st
        self._upload(key, pos)
        return self

    def export(self):
        host = self.state[:_lib.RNG_WORDS].cpu().numpy().view(np.uint32)
        return host[:MT_N].copy(), int(host[MT_N])

    def release_to_torch(self, generator: torch.Generator | None = None):
        key, pos = self.export()
        st = pack_torch_state(self.seed, key, pos, self._template)
        if generator is not None:
            generator.set_state(st)
        else:
            torch.set_rng_state(st)

    # -- draws (all enqueue on the current stream) ---------------------------------------------
    def raw(self, n: int, out=None):
        out = out if out is not None else torch.empty(n, dtype=torch.int32, device=self.device)
        check(lib().mdm_rng_raw(ptr(self.state), ptr(out), n, stream_ptr(self.device)))
        self.words_drawn += n
        return out

    def skip(self, n: int):
        check(lib().mdm_rng_skip(ptr(self.state), n, stream_ptr(self.device)))
        self.words_drawn += n

    def uniform(self, n: int, a=0.0, b=1.0, out=None):
        out = out if out is not None else torch.empty(n, dtype=torch.float32, device=self.device)
        check(lib().mdm_rng_uniform(ptr(self.state), ptr(out), n, a, b, stream_ptr(self.device)))
        self.words_drawn += n
        return out

    def randint(self, lo: int, hi: int, n: int, out=None):
        out = out if out is not None else torch.empty(n, dtype=torch.int64, device=self.device)
        check(lib().mdm_rng_randint(ptr(self.state), ptr(out), n, lo, hi, stream_ptr(self.device)))
        self.words_drawn += n
        return out

    def normal(self, batch: int, per_sample: int, mean=0.0, std=1.0, ratio=None, out=None):
        out = out if out is not None else torch.empty(batch, per_sample, dtype=torch.float32, device=self.device)
        if ratio is not None:
            assert ratio.dtype == torch.float64 and ratio.is_cuda and ratio.numel() == batch
        check(lib().mdm_rng_normal(ptr(self.state), ptr(out), batch, per_sample, mean, std, ptr(ratio),
                                   stream_ptr(self.device)))
        self.words_drawn += batch * per_sample
        return out

    def threshold_mask(self, ratio, batch: int, per_sample: int, ratio2=None, out=None, out2=None):
        assert ratio.dtype == torch.float64 and ratio.is_cuda and ratio.numel() == batch
        out = out if out is not None else torch.empty(batch, per_sample, dtype=torch.uint8, device=self.device)
        if ratio2 is not None:
            assert ratio2.dtype == torch.float64 and ratio2.is_cuda and ratio2.numel() == batch
            out2 = out2 if out2 is not None else torch.empty(batch, per_sample, dtype=torch.uint8, device=self.device)
        check(lib().mdm_rng_threshold_mask(ptr(self.state), ptr(ratio), ptr(out), ptr(ratio2), ptr(out2),
                                           batch, per_sample, stream_ptr(self.device)))
        self.words_drawn += batch * per_sample
        return (out, out2) if ratio2 is not None else out

    def randperm_mask(self, count, batch: int, hw: int, out=None, words_ws=None):
        assert count.dtype == torch.int64 and count.is_cuda and count.numel() == batch
        out = out if out is not None else torch.empty(batch, hw, dtype=torch.uint8, device=self.device)
        if words_ws is None:
            words_ws = torch.empty(batch * (hw - 1), dtype=torch.int32, device=self.device)
        check(lib().mdm_rng_randperm_mask(ptr(self.state), ptr(count), ptr(out), ptr(words_ws), batch, hw,
                                          stream_ptr(self.device)))
        self.words_drawn += batch * (hw - 1)
        return out
