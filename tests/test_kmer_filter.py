"""Presence-filter addressing (sailfish_b200/csrc/kmer_filter.hpp) compiled for the CPU: no false negatives, false-positive rate at
the size index.cu chooses, and the sector locality the scan kernel relies on."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("kf") / "kmer_filter_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", out, os.path.join(ROOT, "tests", "kmer_filter_test.cpp")])
    return out


@pytest.mark.parametrize("k", [31, 21, 19, 15])
def test_filter_addressing(exe, k):
    r = subprocess.run([exe, str(k)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    tag, fpr, sectors = r.stdout.split()
    assert tag == "ok"
    if k == 31:
        assert float(sectors) < 10.0          # 46 successive k-mers of a 76-base read touch ~8 sectors, not 46
