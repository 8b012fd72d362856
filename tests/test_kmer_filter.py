"""The m-mer presence bitmap (sailfish_b200/csrc/kmer_filter.hpp) and the skipping rule of the scan kernel, compiled for the CPU: the
skipping scan finds exactly the positions the one-position-at-a-time scan finds, for dense and sparse bitmaps and several (k, m)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("kf") / "kmer_filter_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", out, os.path.join(ROOT, "tests", "kmer_filter_test.cpp")])
    return out


@pytest.mark.parametrize("k,m,n", [(11, 6, 3000), (11, 6, 300), (13, 7, 20000), (9, 9, 5000), (11, 10, 100000), (15, 8, 40000), (31, 12, 2000000)])
def test_skipping_scan_equals_plain_scan(exe, k, m, n):
    r = subprocess.run([exe, str(k), str(m), str(n)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.split()[0] == "ok"
