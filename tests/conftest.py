import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def sample_data():
    return dict(np.load(os.path.join(GOLDEN, "sample_data.npz")))


@pytest.fixture(scope="session")
def synth_em():
    return dict(np.load(os.path.join(GOLDEN, "synth_em.npz")))


@pytest.fixture(scope="session")
def ctx():
    """One sfb200 context on cuda:0 for the gpu tests (fails loudly without a device: there is no CPU fallback)."""
    from sailfish_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def split_seqs(seq, txp_len):
    off = np.zeros(len(txp_len) + 1, np.int64)
    off[1:] = np.cumsum(txp_len.astype(np.int64))
    return [seq[off[i]:off[i + 1]].tobytes() for i in range(len(txp_len))]
