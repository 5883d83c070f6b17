"""CPU: a model of the operand-ring protocol of the folded-GroupNorm convolution (csrc/igemm.cu, kNorm).

The ring has two slots.  Per work item it is used KC times for halo chunks (filled by the transform warps) and n_sc
times for the patches of the fused 1x1 shortcut (filled by TMA from the A-producer thread); the MMA issuer consumes the
uses in order.  Synchronisation is by mbarriers, whose waits are PARITY waits: `try_wait(parity p)` is true iff the
barrier's count of completed phases c has (c & 1) != p -- a waiter cannot tell phase k from phase k + 2.

Three protocols are simulated under randomly interleaved schedules:

* `final`   -- what the kernel does: each filler has its OWN "slot free" barrier per slot, and the MMA issuer commits
               the consumption of a use to the barrier of whoever fills that slot NEXT (two uses on).  Must be safe
               (no slot is overwritten before its previous use was consumed, nothing is consumed before it is filled)
               and live (every schedule terminates) for every (KC, n_sc, items).
* `skip`    -- first cut: one barrier per slot, a filler simply skips the uses it does not fill.  Its parity tracking
               falls two phases behind and it overwrites a slot that is still being read: the model must find that.
* `observe` -- second cut: one barrier per slot, both fillers WAIT on every use (also the ones they do not fill) to stay
               in phase.  A filler that is busy filling while the other two roles run ahead misses a phase and then waits
               for a parity that never comes back: the model must find that deadlock (the GPU run of round 2 did).
"""
import random

import pytest


class Bar:
    """mbarrier with arrival count 1: `arrive` completes the current phase"""

    def __init__(self):
        self.c = 0

    def arrive(self):
        self.c += 1

    def test(self, parity):
        return (self.c & 1) != parity


def simulate(protocol, KC, n_sc, items, seed, max_steps=200000):
    rnd = random.Random(seed)
    U = KC + n_sc
    uses = [("halo" if (u % U) < KC else "sc") for u in range(items * U)]
    n = len(uses)
    a_full = [Bar(), Bar()]
    a_empty = [Bar(), Bar()]          # `final`: owned by the shortcut filler
    h_empty = [Bar(), Bar()]          # `final`: owned by the transform warps
    filled = [False] * n
    consumed = [False] * n
    errors = []

    def next_is_halo(u):              # who fills this slot next (use u + 2)?
        i = u % U
        return i + 2 < KC or i + 2 >= U

    # ---- roles as generators: yield = a scheduling point; yield "blocked" = could not progress this turn ----------
    def mma():
        pa = [0, 0]
        for u in range(n):
            s = u & 1
            while not a_full[s].test(pa[s]):
                yield "blocked"
            pa[s] ^= 1
            if not filled[u]:
                errors.append(f"use {u} consumed before it was filled")
            for _ in range(rnd.randint(1, 4)):      # the MMAs of this use take a while
                yield
            consumed[u] = True
            if protocol == "final":
                (h_empty if next_is_halo(u) else a_empty)[s].arrive()
            else:
                a_empty[s].arrive()
            yield

    def filler(kind):
        """kind = 'halo' (transform warps) or 'sc' (A-producer thread)"""
        k_own = [0, 0]                # waits done on the own barrier (`final`)
        pa = [0, 0]                   # shared-barrier protocols: parity per slot, advanced per use this role touches
        for u in range(n):
            s = u & 1
            mine = uses[u] == kind
            if protocol == "final":
                if not mine:
                    continue
                bar = h_empty[s] if kind == "halo" else a_empty[s]
                parity = (k_own[s] & 1) ^ 1 if kind == "halo" else (k_own[s] & 1)
                # the transform warps' first use of a slot has no predecessor: parity trick passes on a fresh barrier;
                # a shortcut use always has one inside its own item (KC >= 2)
                while not bar.test(parity):
                    yield "blocked"
                k_own[s] += 1
            elif protocol == "skip":
                if not mine:
                    pa[s] ^= 1        # keeps the slot / parity bookkeeping, does not wait
                    continue
                while not a_empty[s].test(pa[s] ^ 1):
                    yield "blocked"
                pa[s] ^= 1
            elif protocol == "observe":
                while not a_empty[s].test(pa[s] ^ 1):
                    yield "blocked"
                pa[s] ^= 1
                if not mine:
                    yield
                    continue
            # ---- fill use u
            if u >= 2 and not consumed[u - 2]:
                errors.append(f"{kind} filler overwrote slot {s} for use {u} while use {u - 2} was still being read")
            for _ in range(rnd.randint(1, 6) if kind == "halo" else rnd.randint(0, 2)):   # the transform is slow, TMA is not
                yield
            filled[u] = True
            a_full[s].arrive()
            yield

    roles = {"mma": mma(), "halo": filler("halo"), "sc": filler("sc")}
    blocked_streak = 0
    steps = 0
    while roles and steps < max_steps and not errors:
        steps += 1
        name = rnd.choice(sorted(roles))
        try:
            r = next(roles[name])
        except StopIteration:
            del roles[name]
            blocked_streak = 0
            continue
        if r == "blocked":
            blocked_streak += 1
            if blocked_streak > 600:       # every live role polled many times without progress
                return "deadlock", steps
        else:
            blocked_streak = 0
    if errors:
        return errors[0], steps
    if roles:
        return "did not finish", steps
    assert all(consumed)
    return "ok", steps


SHAPES = [(2, 0), (4, 0), (2, 4), (2, 3), (4, 3), (4, 2), (6, 3), (12, 3), (3, 1), (8, 4)]


@pytest.mark.parametrize("KC,n_sc", SHAPES)
def test_final_protocol_is_safe_and_live(KC, n_sc):
    for items in (1, 2, 5):
        for seed in range(40):
            res, _ = simulate("final", KC, n_sc, items, seed)
            assert res == "ok", (KC, n_sc, items, seed, res)


def test_skipping_other_fillers_uses_aliases_the_parity():
    """the first cut is caught: with a fused shortcut and several items some schedule overwrites a live slot"""
    found = set()
    for KC, n_sc in [(4, 3), (2, 3), (4, 2)]:
        for seed in range(200):
            res, _ = simulate("skip", KC, n_sc, 4, seed)
            if res != "ok":
                found.add(res.split(" ")[0] if res != "deadlock" else res)
    assert found, "the model should expose the two-phase aliasing of the skip protocol"


def test_observer_waits_deadlock_when_a_role_misses_a_phase():
    """the second cut is caught too: a role that is busy (or simply not scheduled) while the other two run ahead by two
    phases waits for a parity that never comes back.  On the GPU the shortcut shapes hit it (the transform warps are busy
    for thousands of cycles per chunk); the model, whose scheduler may starve any role, shows that a pure observer is
    formally unsafe even without a shortcut."""
    for KC, n_sc in [(4, 3), (4, 0)]:
        bad = 0
        for seed in range(300):
            res, _ = simulate("observe", KC, n_sc, 4, seed)
            bad += res != "ok"
        assert bad > 0, "the model should expose the missed-phase deadlock of the observer protocol"


def test_single_chunk_layers_with_a_shortcut_are_outside_the_protocol():
    """why mdm_conv_fprop refuses gn_coef for cin < 128 (csrc/igemm.cu: `cin >= 128`): with ONE halo chunk per item and a
    shortcut, "the use two on" of an item's last uses is no longer guaranteed to be a halo chunk of the next item, the
    MMA issuer's rule picks the wrong barrier and the model deadlocks or overwrites a live slot; without a shortcut a
    single chunk is fine."""
    for n_sc in (1, 2, 3):
        assert any(simulate("final", 1, n_sc, 4, seed)[0] != "ok" for seed in range(50))
    assert all(simulate("final", 1, 0, 4, seed)[0] == "ok" for seed in range(50))
