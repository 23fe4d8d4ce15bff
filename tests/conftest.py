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


def golden_cases(prefix):
    return sorted(f[len(prefix) + 1:-4] for f in os.listdir(GOLDEN) if f.startswith(prefix + "_") and f.endswith(".npz"))


def load_golden(prefix, name):
    z = np.load(os.path.join(GOLDEN, "%s_%s.npz" % (prefix, name)))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
