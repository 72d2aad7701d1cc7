import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "hostsim"), os.path.join(ROOT, "oracle", "refshim")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_files():
    return sorted(f for f in glob.glob(os.path.join(GOLDEN_DIR, "*_[0-9][0-9][0-9].npz")) if not os.path.basename(f).startswith("ppo_"))


def load_golden(path):
    z = np.load(path)
    g = {k: z[k] for k in z.files}
    g["variant"] = str(g["variant"])
    return g


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O
    O.build()
    return O
