"""CPU test: the host-side scheduling of an optimizer run that stops at iterations 50 / 500 / 1000 for the bias / GC effective-length
recomputation (sailfish_b200/csrc/em_segments.hpp, used by sfb200_em_run_bias) -- see tests/em_segments_test.cpp."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_em_segments_equal_one_loop_with_a_hook(tmp_path):
    exe = str(tmp_path / "em_segments_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "em_segments_test.cpp")])
    assert "em_segments: ok" in subprocess.check_output([exe]).decode()
