import glob
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


def golden_files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, f"{prefix}_*.npz")))


def load_golden(path):
    d = np.load(path, allow_pickle=False)
    out = {k: d[k] for k in d.files}
    for k in ("env_id", "mode"):
        if k in out:
            out[k] = str(out[k])
    for k in ("seed", "width", "height", "max_steps"):
        if k in out:
            out[k] = int(out[k])
    if "see_through" in out:
        out["see_through"] = bool(out["see_through"])
    return out


@pytest.fixture(scope="session")
def has_cuda():
    import torch

    return torch.cuda.is_available()
