"""CPU: the C-ABI library loads and exports every symbol include/mdm.h declares; the ctypes
table in mdm_b200/_lib.py covers exactly those; compute entry points refuse to run without CUDA."""
import ctypes
import re

import numpy as np
import pytest
import torch

from mdm_b200 import _lib


def header_symbols():
    src = open(_lib.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mdm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    l = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(l, s), f"{s} declared in mdm.h but not exported"


def test_ctypes_table_matches_header():
    import importlib
    for mod in ("mdm_b200.denoiser_ops", "mdm_b200.optim_ops", "mdm_b200.comm_ops"):   # register their entry points
        try:
            importlib.import_module(mod)
        except ModuleNotFoundError:
            pass
    assert sorted(_lib._SIGS) == header_symbols()


def test_seed_host_matches_torch():
    host = np.zeros(_lib.RNG_WORDS, dtype=np.uint32)
    _lib.check(_lib.lib().mdm_rng_seed_host(host.ctypes.data, 1234))
    torch.manual_seed(1234)
    from mdm_b200.rng import parse_torch_state
    _, key, pos = parse_torch_state(torch.get_rng_state())
    assert np.array_equal(host[:624], key) and host[624] == pos == 624


def test_no_cpu_fallback():
    import scheduler
    from oracle.mdm_oracle import default_args
    S = scheduler.Scheduler(default_args(data_size=8, ddpm_num_steps=10))
    S.update_ddpm_num_steps(10)
    x = torch.zeros(2, 3, 8, 8)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            S.degrade_training(torch.tensor([0.5, 0.5], dtype=torch.float64), x, "degraded_area", "image-wise")
    assert _lib.lib().mdm_version() >= 100


def test_reserve_sms_is_host_state():
    """mdm_reserve_sms only records how many SMs later persistent GEMM launches leave free (host side, no GPU needed):
    set / query / clamp / restore"""
    l = _lib.lib()
    before = l.mdm_reserve_sms(3)
    try:
        assert l.mdm_reserve_sms(-1) == 3            # query
        assert l.mdm_reserve_sms(10_000) == 3        # clamped to 147 of the 148 SMs
        assert l.mdm_reserve_sms(-1) == 147
    finally:
        l.mdm_reserve_sms(before)
    assert l.mdm_reserve_sms(-1) == before
