"""Multi-rank parity on real GPUs (-m gpu; skipped on a box with fewer than two): tests/multi_gpu_check.py under torchrun on 2 ranks --
every rank maps its shard of one read set, and the merged-classes EM (default) and the per-iteration all-reduce EM
(SFB200_MULTI_EM_ALLREDUCE=1) must both reproduce the single-process oracle on all reads (classes, counters, the ordered
fragment-length sample, EM and VBEM estimates)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.gpu
def test_two_ranks_match_single_process_oracle():
    if n_gpus() < 2:
        pytest.skip("needs two GPUs (the driver's scaling run covers N = 2, 4, 8)")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "multi-GPU check ok on 2 ranks" in r.stdout
