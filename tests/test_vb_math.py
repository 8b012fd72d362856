"""sfb_digamma / sfb_exp_digamma (sailfish_b200/csrc/vb_math.hpp, what the VBEM kernels evaluate per transcript and iteration), compiled
for the CPU and checked against mpmath at 50 digits: VBEM's expTheta = exp(digamma(alpha) - digamma(sum alpha))
(/root/reference/src/CollapsedEMOptimizer.cpp:398-416) is computed as sfb_exp_digamma(alpha) * exp(-digamma(sum alpha))."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("vb") / "vb_math_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-ffp-contract=off", "-o", out, os.path.join(ROOT, "tests", "vb_math_test.cpp")])
    return out


def run(exe, xs):
    r = subprocess.run([exe], input="\n".join("%.17g" % x for x in xs) + "\n", capture_output=True, text=True, check=True)
    a = np.array([[float(v) for v in line.split()] for line in r.stdout.strip().split("\n")])
    return a[:, 0], a[:, 1]


def test_exp_digamma_against_mpmath(exe):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    rng = np.random.default_rng(0)
    xs = np.concatenate([10 ** rng.uniform(-3, 1.3, 3000), rng.uniform(10, 40, 1500), 10 ** rng.uniform(1.3, 9, 1500),
                         [1e-3, 0.999999, 1.0, 15.999999, 16.0, 16.000001, 1e12]])
    dg, edg = run(exe, xs)
    worst_e = worst_series = 0.0
    for x, d, e in zip(xs, dg, edg):
        psi = mp.digamma(mp.mpf(float(x)))
        ref = mp.exp(psi)
        if ref > mp.mpf("1e-290"):
            err = float(abs(mp.mpf(float(e)) - ref) / ref)
            worst_e = max(worst_e, err)
            if x >= 16:
                worst_series = max(worst_series, err)
        # digamma itself: absolute error near its root at 1.4616, relative elsewhere
        assert float(abs(mp.mpf(float(d)) - psi)) <= 4e-15 * max(1.0, float(abs(psi)))
    assert worst_series < 5e-16               # the twelve-term series alone
    assert worst_e < 3e-13                    # tiny alphas: exp(-1/x) carries |1/x| ulps


def test_exp_digamma_underflows_like_the_reference_form(exe):
    """alpha = the prior of a transcript without reads (1e-3 and below): exp(digamma(alpha) - logNorm) is exactly 0 in the reference
    (digamma ~ -1000), and so is the product form"""
    xs = np.array([1e-3, 1.2e-3, 5e-4, 1e-8])
    _, edg = run(exe, xs)
    assert (edg == 0.0).all()
    for x in (2e-3, 1e-2, 0.5, 3.0, 100.0, 1e7):
        d, e = run(exe, [x])
        assert e[0] > 0 and abs(np.log(e[0]) - d[0]) <= 1e-12 * max(1.0, abs(d[0]))
