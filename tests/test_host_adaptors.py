"""The C++ adaptors (sailfish_b200/host/sfb200_host.hpp: same class / method names as the reference) compile with plain g++
against the C ABI (CPU test) and, on a GPU, reproduce the oracle through the whole mainQuantify call sequence."""
import os
import subprocess

import numpy as np
import pytest

from conftest import split_seqs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "_host_adaptor_test")


def build_exe():
    from sailfish_b200 import capi
    capi.lib()
    lib_dir = os.path.join(ROOT, "sailfish_b200")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-o", EXE, os.path.join(ROOT, "tests", "host_adaptor_test.cpp"),
                           "-L" + lib_dir, "-lsfb200", "-Wl,-rpath," + lib_dir, "-pthread"])


def test_adaptors_compile_with_plain_gxx():
    build_exe()
    assert os.path.exists(EXE)
    # without a GPU the Device constructor must fail loudly (exit code 3 from the sfb200::Error handler)
    import torch
    if not torch.cuda.is_available():
        tmp = os.path.join(ROOT, "tests", "_host_in.txt")
        open(tmp, "w").write("1\n100 ACGTACGTACGTACGTACGTACGTACGTACGTACGT\n0\n")
        rc = subprocess.call([EXE, tmp], stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL)
        os.remove(tmp)
        assert rc == 3


def test_bias_model_cdf_equals_empirical_distribution(tmp_path):
    from oracle import pyoracle as O
    from sailfish_b200 import capi
    O.lib(); capi.lib()
    exe = str(tmp_path / "bias_model_test")
    odir, ldir = os.path.join(ROOT, "oracle"), os.path.join(ROOT, "sailfish_b200")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "bias_model_test.cpp"),
                           "-L" + odir, "-loracle", "-Wl,-rpath," + odir, "-L" + ldir, "-lsfb200", "-Wl,-rpath," + ldir, "-pthread"])
    assert "bias model ok" in subprocess.check_output([exe]).decode()


@pytest.mark.gpu
def test_adaptors_reproduce_oracle(sample_data, tmp_path):
    from oracle import pyoracle as O
    d = sample_data
    build_exe()
    seqs = split_seqs(d["txp_seq"], d["txp_len"])
    r1 = [d["reads1"][int(d["off1"][i]):int(d["off1"][i + 1])].tobytes().decode() for i in range(len(d["off1"]) - 1)]
    r2 = [d["reads2"][int(d["off2"][i]):int(d["off2"][i + 1])].tobytes().decode() for i in range(len(d["off2"]) - 1)]
    inp = tmp_path / "in.txt"
    with open(inp, "w") as f:
        f.write("%d\n" % len(seqs))
        for s, e in zip(seqs, d["eff"]):
            f.write("%.17g %s\n" % (e, s.decode()))
        f.write("%d\n" % len(r1))
        for a, b in zip(r1, r2):
            f.write("%s %s\n" % (a, b))
    out = subprocess.check_output([EXE, str(inp)]).decode().split("\n")
    counters = [int(x) for x in out[0].split()]
    assert counters == [int(x) for x in d["counters"]]
    n_cls, tot = (int(x) for x in out[1].split())
    assert n_cls == len(d["counts"]) and tot == int(d["counts"].sum())
    T = len(seqs)
    est = np.array([float(out[2 + t].split()[0]) for t in range(T)])
    mass = np.array([float(out[2 + t].split()[1]) for t in range(T)])
    np.testing.assert_allclose(est, d["ref_est_vb0"], rtol=1e-4, atol=1e-6)       # the reference optimizer's own estimates
    np.testing.assert_allclose(mass, d["ref_mass_vb0"], rtol=1e-4, atol=1e-9)
    boots = [float(l.split()[1]) for l in out if l.startswith("boot")]
    gibbs = [int(l.split()[1]) for l in out if l.startswith("gibbs")]
    assert len(boots) == 3 and all(abs(b - tot) < 1e-6 * tot for b in boots)
    assert len(gibbs) == 3 and all(g == tot for g in gibbs)
