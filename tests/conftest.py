import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        with np.load(GOLDEN / name) as d:
            return {k: d[k] for k in d.files}
    return load


@pytest.fixture(scope="session")
def oracle():
    from oracle import ganspace_oracle
    return ganspace_oracle


@pytest.fixture(scope="session")
def mapping_weights(oracle):
    ws, bs = oracle.mapping_random_init(1234)
    return ws, bs
