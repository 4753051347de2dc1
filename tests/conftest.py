import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))


@pytest.fixture(scope="session")
def golden_preview():
    """outputs of the reference's ray_marcher (pathtracer.py:471-685), tests/golden/gen_golden_preview.py"""
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_preview_v1.npz"))


@pytest.fixture(scope="session")
def luts():
    from oracle import oracle as orc
    return orc.load_luts()
