"""CPU test: the per-position arithmetic the bias / GC kernels run (sailfish_b200/csrc/bias_core.inl, compiled here as host
code) replayed serially against the pinned CPU oracle -- see tests/bias_core_test.cpp.  The CUDA launch code around it
(sailfish_b200/csrc/bias.cu) is covered by tests/test_gpu_bias.py."""
import os
import subprocess

from oracle import pyoracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bias_core_matches_oracle(tmp_path):
    O.lib()                                                           # builds oracle/liboracle.so if needed
    exe = str(tmp_path / "bias_core_test")
    odir = os.path.join(ROOT, "oracle")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "bias_core_test.cpp"),
                           "-L" + odir, "-loracle", "-Wl,-rpath," + odir])
    assert "bias core ok" in subprocess.check_output([exe]).decode()
