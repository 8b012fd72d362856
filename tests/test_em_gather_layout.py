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


def test_dense_layout_and_iteration(tmp_path):
    """the dense-component layout (sailfish_b200/csrc/em_dense_build.inl): components, slots, masks, and the per-component
    iteration of k_em_dense against the reference-shaped update -- see tests/em_dense_layout_test.cpp"""
    exe = str(tmp_path / "em_dense_layout_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "em_dense_layout_test.cpp")])
    out = subprocess.check_output([exe]).decode()
    assert "em_dense layout ok" in out
