"""A model of k_em_dense's lagged stopping rule (sailfish_b200/csrc/em_dense.cuh): CTAs that run at random relative speeds publish
their part of iteration m into slot m mod DN_LAG_SLOTS, send the arrival for m one iteration later, read what the grid found in
iteration m - DN_LAG (a load before the sweep, a wait after it if that came too early), and CTA 0 clears the slot of
m + DN_LAG + 1.  The model checks what the kernel's comments claim: a slot is only ever read when it holds exactly the
contributions of the iteration asked for, from every CTA; nobody publishes into a slot before it was cleared of its previous use;
every CTA stops at the same iteration, the one a synchronous loop would stop at; and the alpha ring (DN_LAG + 1 entries) still holds
that iteration.  The constants are read from the header, so changing them there re-runs the argument."""
import os
import random
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "sailfish_b200", "csrc", "em_dense.cuh")).read()
LAG, RING, SLOTS = (int(re.search(r"%s = (\d+)" % n, SRC).group(1)) for n in ("DN_LAG", "DN_RING", "DN_LAG_SLOTS"))


class Slot:
    def __init__(self):
        self.contrib = []          # (cta, iteration) of every publication since the last clear
        self.arrived = 0           # grows for ever, like the kernel's counter


def run_model(n_cta, rel, min_iter, max_iter, tol, seed, lag=None, slots=None, ring=None):
    """rel(cta, m) -> this CTA's max relative change in iteration m.  Returns the iteration every CTA stopped at."""
    lag = LAG if lag is None else lag
    slots_n = SLOTS if slots is None else slots
    ring = RING if ring is None else ring
    rng = random.Random(seed)
    S = [Slot() for _ in range(slots_n)]
    want = lambda x: n_cta * ((x - 1) // slots_n + 1)
    state = [dict(m=0, pc="top", pf=None, stop=None, consumed=0) for _ in range(n_cta)]
    published = [0] * n_cta        # the last iteration a CTA has published

    def read_slot(x):
        s = S[x % slots_n]
        assert s.arrived >= want(x)
        its = {it for _, it in s.contrib}
        assert its <= {x}, "slot of iteration %d holds contributions of %s" % (x, sorted(its))
        assert len(s.contrib) == n_cta, "slot of iteration %d read with %d of %d contributions" % (x, len(s.contrib), n_cta)
        return max(v for v, _ in [(rel(c, x), c) for c, _ in s.contrib])

    def decide(c, x, st):
        mr = read_slot(x)
        st["consumed"] = x
        if x >= min_iter and (x >= max_iter or not (mr > tol)):
            assert st["m"] - x < ring, "alpha of iteration %d has left the ring at iteration %d" % (x, st["m"])
            st["stop"] = x
            return True
        return False

    steps = 0
    while any(st["stop"] is None for st in state):
        steps += 1
        assert steps < 10_000_000, "deadlock"
        c = rng.randrange(n_cta)
        st = state[c]
        if st["stop"] is not None:
            continue
        if st["pc"] == "top":                                   # loop top: the load of what is due, then the sweep
            st["m"] += 1
            m = st["m"]
            x = m - lag
            st["pf"] = (S[x % slots_n].arrived >= want(x)) if x >= 1 else None
            st["pc"] = "post"
        elif st["pc"] == "post":                                # after the CTA barrier: thread 0's work
            m = st["m"]
            if m > 1:
                S[(m - 1) % slots_n].arrived += 1               # the arrival of the iteration before
            s = S[m % slots_n]
            assert all(it == m for _, it in s.contrib), "CTA %d publishes iteration %d into a slot that still holds %s" % (
                c, m, sorted({it for _, it in s.contrib}))
            s.contrib.append((c, m))
            published[c] = m
            if c == 0:
                z = m + lag + 1
                zs = S[z % slots_n]
                assert all(st2["consumed"] >= z - slots_n or z - slots_n < 1 for st2 in state), "slot cleared before everybody read it"
                assert not any(it == z for _, it in zs.contrib), "slot cleared after somebody published iteration %d into it" % z
                zs.contrib = []
            final = m >= max_iter and m >= min_iter
            if final:
                S[m % slots_n].arrived += 1
            st["final"] = final
            st["pc"] = "consume"
            st["x"] = max(1, m - lag) if final else m - lag
        elif st["pc"] == "consume":
            m, x = st["m"], st["x"]
            if x < 1:
                st["pc"] = "top"
                continue
            if S[x % slots_n].arrived < want(x):
                continue                                        # spinning
            if decide(c, x, st):
                continue
            if st["final"] and x < m:
                st["x"] = x + 1
            else:
                st["pc"] = "top"
    stops = {st["stop"] for st in state}
    assert len(stops) == 1, stops
    return stops.pop()


def sync_stop(n_cta, rel, min_iter, max_iter, tol):
    m = 0
    while True:
        m += 1
        mr = max(rel(c, m) for c in range(n_cta))
        if m >= min_iter and (m >= max_iter or not (mr > tol)):
            return m


@pytest.mark.parametrize("seed", range(12))
def test_lagged_rule_equals_synchronous_rule(seed):
    rng = random.Random(1000 + seed)
    n_cta = rng.choice([1, 2, 7, 32])
    decay = rng.uniform(0.6, 0.97)
    bump = {(rng.randrange(n_cta), rng.randrange(1, 60)) for _ in range(5)}      # a CTA that is not converged a little longer
    rel = lambda c, m: decay ** m * (1.0 + 0.1 * c / n_cta) + (0.5 if (c, m) in bump else 0.0)
    for min_iter, max_iter, tol in ((1, 10_000, 0.01), (50, 10_000, 0.01), (1, rng.choice([1, 2, 3, 4, 5, 17]), 1e-9), (40, 30, 0.5), (1, 10_000, 0.9)):
        want = sync_stop(n_cta, rel, min_iter, max_iter, tol)
        assert run_model(n_cta, rel, min_iter, max_iter, tol, seed) == want


def test_model_catches_a_slot_ring_that_is_too_small():
    """with fewer slots than 3 * DN_LAG + 2 the argument for clearing the slot of m + DN_LAG + 1 no longer holds: the model must say so
    for some interleaving (this is what makes the test above mean something)"""
    rel = lambda c, m: 0.97 ** m
    caught = 0
    for seed in range(40):
        try:
            run_model(8, rel, 1, 10_000, 0.01, seed, slots=LAG + 2)
        except AssertionError:
            caught += 1
    assert caught > 0
    assert SLOTS > 3 * LAG + 1 and RING == LAG + 1
