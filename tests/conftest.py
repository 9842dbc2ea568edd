import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle

    oracle.build()
    oracle.set_num_threads(min(os.cpu_count() or 1, 32))
    return oracle


@pytest.fixture(scope="session")
def ctx():
    """The product path: the C ABI over the CUDA kernels.  Fails loudly when the library/GPU is missing."""
    import sfm_gms_b200 as sg

    c = sg.Context(0)
    yield c
    c.close()
