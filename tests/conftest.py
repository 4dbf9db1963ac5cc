import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def emu():
    from tests.emu import emu as E
    E.build()
    return E


@pytest.fixture(scope="session")
def golden_root():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "nlp_root.npz"))


@pytest.fixture(scope="session")
def golden_gym():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "nlp_gym.npz"))
