"""CPU: the GF(2) jump-ahead arithmetic behind the multi-CTA mt19937 stream (csrc/mt_jump.cu) against numpy's MT19937
(the same engine torch's CPU generator is: SURVEY.md section 3.2.1).  Host-only entry points of the C ABI -- the
characteristic polynomial comes from Berlekamp-Massey inside the library; nothing here is tabulated.

  * mdm_rng_advance_host: the state after n draws equals numpy's state after random_raw(n), for short hops (block
    stepping) and far jumps (polynomial method), from seeded and mid-block positions;
  * mdm_rng_jump_table_host: polynomial c applied to the word sequence the way the device kernel does it
    (y[J + m] = XOR_{i : g_i} y[i + m] over windows starting at y[1]) reproduces the window (c * bpc - 1) blocks ahead."""
import numpy as np
import pytest

from mdm_b200 import _lib


def np_engine(seed):
    bg = np.random.MT19937()
    bg._legacy_seeding(seed)
    return bg


def state_of(bg):
    st = bg.state["state"]
    out = np.empty(625, dtype=np.uint32)
    out[:624] = st["key"]
    out[624] = st["pos"]
    return out


@pytest.mark.parametrize("seed,pre,n", [(0, 0, 1), (0, 0, 624), (0, 0, 625), (1234, 7, 623), (1234, 7, 624 * 40 + 1),
                                        (5, 300, 624 * 41), (5, 300, 4194304), (77, 623, 12582912 + 3), (3, 1, 100_000_007)])
def test_advance_host_matches_numpy(seed, pre, n):
    bg = np_engine(seed)
    if pre:
        bg.random_raw(pre)
    s_in = state_of(bg)
    out = np.zeros(625, dtype=np.uint32)
    _lib.check(_lib.lib().mdm_rng_advance_host(s_in.ctypes.data, n, out.ctypes.data))
    left = n
    while left > 0:                        # numpy draws in chunks (memory)
        m = min(left, 1 << 24)
        bg.random_raw(m)
        left -= m
    want = state_of(bg)
    assert out[624] == want[624], (out[624], want[624])
    assert np.array_equal(out[:624], want[:624])


def untempered(seed, pre, count):
    """y[0 .. count): untempered words continuing the state (key words from the current block on)"""
    bg = np_engine(seed)
    if pre:
        bg.random_raw(pre)
    key = state_of(bg)[:624].astype(np.uint64)
    y = np.empty(count, dtype=np.uint64)
    y[:624] = key
    for k in range(624, count):            # plain recurrence (scalar loop: the checker, not the product)
        u = (y[k - 624] & 0x80000000) | (y[k - 623] & 0x7fffffff)
        y[k] = y[k - 227] ^ (u >> 1) ^ (0x9908b0df if (y[k - 623] & 1) else 0)
    return y.astype(np.uint32)


def test_jump_table_polynomials_reproduce_far_windows():
    n_polys, bpc = 3, 40
    polys = np.zeros(n_polys * 624, dtype=np.uint32)
    _lib.check(_lib.lib().mdm_rng_jump_table_host(polys.ctypes.data, n_polys, bpc))
    need = (n_polys * bpc) * 624 + 19968 + 700
    y = untempered(11, 5, need)
    z = y[1:]                                # windows start at y[1]
    for c in range(1, n_polys + 1):
        bits = np.unpackbits(polys[(c - 1) * 624:c * 624].view(np.uint8), bitorder="little")
        assert bits[19937:].sum() == 0 and bits.sum() > 1          # reduced: degree < 19937 (sparse for near jumps, ~10^4 terms far out)
        win = np.zeros(624, dtype=np.uint32)
        for i in np.nonzero(bits)[0]:
            win ^= z[i:i + 624]
        J = (c * bpc - 1) * 624
        assert np.array_equal(win, z[J:J + 624]), c
