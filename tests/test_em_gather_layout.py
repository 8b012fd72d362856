"""CPU test: the per-CTA gather layout of the atomic-free EM loop (sailfish_b200/csrc/em_gather_build.inl, compiled as a
single-thread host function) is a correct transpose pair, and the beta / r iteration over it reproduces the update written in
the reference's shape (CollapsedEMOptimizer.cpp:235-277, :760-769) -- see tests/em_gather_layout_test.cpp."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gather_layout_and_iteration(tmp_path):
    exe = str(tmp_path / "em_gather_layout_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "em_gather_layout_test.cpp")])
    out = subprocess.check_output([exe]).decode()
    assert "em_gather layout ok" in out
