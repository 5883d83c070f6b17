"""GPU: the device mt19937 stream (csrc/rng.cu) against torch-generated known answers and the
oracle -- bit-exact words, uniforms, permutations, thresholds and generator state."""
import numpy as np
import pytest
import torch

from oracle.mdm_oracle import OracleRNG
from tests.helpers import torch_state_words

pytestmark = pytest.mark.gpu


def dev_rng(seed):
    from mdm_b200.rng import DeviceMT19937
    return DeviceMT19937("cuda").manual_seed(seed)


@pytest.mark.parametrize("seed", [0, 1234])
def test_known_answers(golden, seed):
    g = golden("rng_kat")
    u32 = lambda t: t.cpu().numpy().view(np.uint32)
    assert np.array_equal(u32(dev_rng(seed).raw(2000)), g[f"raw_{seed}"])
    assert np.array_equal(dev_rng(seed).uniform(700, 0.0, 1.0).cpu().numpy(), g[f"rand_{seed}"])
    assert np.array_equal(dev_rng(seed).uniform(700, -1.0, 1.0).cpu().numpy(), g[f"uniform_{seed}"])
    assert np.array_equal(dev_rng(seed).randint(0, 394, 700).cpu().numpy(), g[f"randint394_{seed}"])
    n = dev_rng(seed).normal(1, 1280, 0.5, 2.0).cpu().numpy().reshape(-1)
    np.testing.assert_allclose(n, g[f"normal_{seed}"], atol=4e-6, rtol=0)   # 2 ulp at |x| <= 12 (std 2)
    r = dev_rng(seed)
    r.raw(1000)
    key, pos = r.export()
    gk, gp = torch_state_words(g[f"state_after_1000_{seed}"])
    assert pos == gp and np.array_equal(key, gk)


def test_split_draws_continue_the_stream():
    r = dev_rng(5)
    parts = [r.raw(n).cpu().numpy().view(np.uint32) for n in (1, 623, 1, 624, 625, 7, 3000)]
    assert np.array_equal(np.concatenate(parts), OracleRNG(5).raw(sum((1, 623, 1, 624, 625, 7, 3000))))
    r.skip(1234)
    o = OracleRNG(5)
    o.raw(4881 + 1234)
    assert np.array_equal(r.raw(10).cpu().numpy().view(np.uint32), o.raw(10))


def test_adopt_and_release_torch_generator():
    from mdm_b200.rng import DeviceMT19937
    torch.manual_seed(99)
    torch.rand(777)
    r = DeviceMT19937("cuda").adopt_torch()
    a = r.uniform(5000, 0.0, 1.0).cpu()
    b = torch.rand(5000)
    assert torch.equal(a, b)
    r.release_to_torch()
    # torch continues where the device stream stopped == where torch itself stopped
    c = torch.rand(100)
    torch.manual_seed(99)
    torch.rand(777 + 5000)
    assert torch.equal(c, torch.rand(100))


@pytest.mark.parametrize("hw,B", [(64, 3), (1024, 8), (4096, 4)])
def test_randperm_mask_bit_exact(hw, B):
    torch.manual_seed(3)
    counts = torch.randint(0, hw + 1, (B,))
    counts[0] = hw
    counts[-1] = 1
    ref = torch.ones(B, hw)
    torch.manual_seed(17)
    for i in range(B):
        ref[i, torch.randperm(hw)[: counts[i]]] = 0.0
    state_after = torch.get_rng_state()
    r = dev_rng(17)
    m = r.randperm_mask(counts.cuda(), B, hw)
    assert torch.equal(m.cpu().float(), ref)
    key, pos = r.export()
    gk, gp = torch_state_words(state_after.numpy())
    assert pos == gp and np.array_equal(key, gk)


@pytest.mark.parametrize("per,B", [(256, 5), (3 * 1024, 4)])
def test_threshold_mask_bit_exact(per, B):
    ratio = torch.tensor([0.001, 0.25, 0.5, 0.999, 1.0][:B], dtype=torch.float64)
    torch.manual_seed(8)
    u = torch.FloatTensor(B, per).uniform_(0.0, 1.0)
    ref = (u > ratio.unsqueeze(1)).float()
    ref2 = (u > (ratio * 0.5).unsqueeze(1)).float()
    m = dev_rng(8).threshold_mask(ratio.cuda(), B, per)
    assert torch.equal(m.cpu().float(), ref)
    m1, m2 = dev_rng(8).threshold_mask(ratio.cuda(), B, per, ratio2=(ratio * 0.5).cuda())
    assert torch.equal(m1.cpu().float(), ref) and torch.equal(m2.cpu().float(), ref2)


def test_large_stream_matches_oracle():
    n = 3_000_000
    a = dev_rng(2).raw(n).cpu().numpy().view(np.uint32)
    assert np.array_equal(a, OracleRNG(2).raw(n))


# ---- multi-CTA generation (GF(2) jump-ahead, csrc/mt_jump.cu + the kernel in csrc/rng.cu) -------------------------------
W = 64 * 624          # words per CTA (mdm_b200.rng.JUMP_BLOCKS_PER_CTA)


@pytest.mark.parametrize("pre,n", [(0, 262144), (1, 262144 + 13), (623, 7 * W), (624, 7 * W + 1), (300, 7 * W - 1),
                                   (17, 12_582_912), (5, 20 * W + 623), (5, 20 * W + 624), (5, 20 * W + 625)])
def test_parallel_stream_is_the_serial_stream(pre, n):
    """draws above MDM_RNG_PAR_MIN_WORDS are cut over ceil(n / W) CTAs: same words, same advanced state, and the next
    (serial) draw continues the sequence"""
    r = dev_rng(31)
    o = OracleRNG(31)
    if pre:
        assert np.array_equal(r.raw(pre).cpu().numpy().view(np.uint32), o.raw(pre))
    got = r.raw(n).cpu().numpy().view(np.uint32)
    want = o.raw(n)
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, (bad[:5], bad.size)
    key, pos = r.export()
    okey, opos = o.state_words()
    assert pos == opos and np.array_equal(key, okey)
    assert np.array_equal(r.raw(1000).cpu().numpy().view(np.uint32), o.raw(1000))
    assert np.array_equal(r.raw(n).cpu().numpy().view(np.uint32), o.raw(n))      # a second parallel draw from a mid-block position


def test_parallel_draw_longer_than_the_table(monkeypatch):
    """a draw that needs more CTAs than there are polynomials runs as several launches, each committing its state"""
    from mdm_b200 import rng as rng_mod
    from mdm_b200._lib import check, lib
    r = dev_rng(9)
    table = next(iter(rng_mod._jump_tables.values()))
    check(lib().mdm_rng_enable_parallel(table.data_ptr(), 4, rng_mod.JUMP_BLOCKS_PER_CTA))     # 5 CTAs per launch
    try:
        n = 12 * W + 5
        got = r.raw(n).cpu().numpy().view(np.uint32)
        o = OracleRNG(9)
        assert np.array_equal(got, o.raw(n))
        key, pos = r.export()
        okey, opos = o.state_words()
        assert pos == opos and np.array_equal(key, okey)
    finally:
        check(lib().mdm_rng_enable_parallel(table.data_ptr(), rng_mod.JUMP_POLYS, rng_mod.JUMP_BLOCKS_PER_CTA))


def test_parallel_masks_and_noise_equal_serial():
    """the consumers of the stream at the configs[3] sizes: threshold masks, uniforms and Box-Muller noise are identical
    with the table lent and with it withdrawn (one CTA), and so is the state left behind"""
    from mdm_b200 import rng as rng_mod
    from mdm_b200._lib import check, lib
    B, hw = 32, 128 * 128
    ratio = torch.linspace(0.01, 0.99, B, dtype=torch.float64).cuda()
    outs = {}
    table = next(iter(rng_mod._jump_tables.values())) if rng_mod._jump_tables else None
    for par in (True, False):
        r = dev_rng(4)
        if not par:
            check(lib().mdm_rng_enable_parallel(None, 0, 0))
        try:
            r.raw(5)
            m = r.threshold_mask(ratio, B, hw)
            m1, m2 = r.threshold_mask(ratio, B, hw, ratio2=ratio * 0.5)
            u = r.uniform(B * hw, -1.0, 1.0)
            nz = r.normal(B, 3 * hw, 0.25, 1.0, ratio=ratio)
            outs[par] = (m.clone(), m1.clone(), m2.clone(), u.clone(), nz.clone(), r.export())
        finally:
            if not par and table is not None:
                check(lib().mdm_rng_enable_parallel(table.data_ptr(), rng_mod.JUMP_POLYS, rng_mod.JUMP_BLOCKS_PER_CTA))
    for a, b in zip(outs[True][:5], outs[False][:5]):
        assert torch.equal(a, b)
    assert outs[True][5][1] == outs[False][5][1] and np.array_equal(outs[True][5][0], outs[False][5][0])
