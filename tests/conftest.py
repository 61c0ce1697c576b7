import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with `-m gpu`")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


GOLDEN_CASES = ["thin_nonoise_jvla1000", "thin_noise0_jvla1000_k100", "mixed_noise1_jvla1000_k30",
                "mixed_noise1_jvla200_k60", "thin_noise2_meerkat300_k40"]
SMALL_CASES = ["mixed_noise1_jvla200_k60", "thin_noise2_meerkat300_k40"]


@pytest.fixture
def meerkat_dataset():
    # same fixture as the reference's tests/conftest.py:5-8
    yield np.linspace(start=0.9e9, stop=1.67e9, num=1000)


@pytest.fixture
def jvla_dataset():
    # same fixture as the reference's tests/conftest.py:11-14
    yield np.linspace(start=1.008e9, stop=2.031e9, num=1000)


@pytest.fixture(scope="session")
def built_library():
    import __graft_entry__ as ge
    ge.build()
    from csromer_b200 import _lib
    return _lib.load()
