import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope='session')
def golden():
    """Golden vectors extracted from the reference's shipped model (tests/golden/make_golden.py)."""
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'golden_pt.npz'))


@pytest.fixture(scope='session')
def oracle():
    """The compiled reference C++ (oracle/_ref/libgcl_ref.so)."""
    from oracle import gcl_ref
    if not gcl_ref.available():
        from pyrsd_b200 import build
        build.build_oracle()
    if not gcl_ref.available():
        pytest.skip("oracle/_ref/libgcl_ref.so not built and /root/reference absent")
    return gcl_ref


@pytest.fixture(scope='session')
def ref_cosmo(oracle, golden):
    return oracle.golden_cosmology(golden)


@pytest.fixture(scope='session')
def ref_plin(oracle, ref_cosmo):
    return oracle.RefLinearPS(ref_cosmo)


@pytest.fixture(scope='session')
def tight():
    """Outputs of the reference C++ run at tight epsrel (tests/golden/make_tight.py)."""
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'oracle_tight.npz'))


@pytest.fixture(scope='session')
def emul():
    """CPU emulation of the host logic (test-only library)."""
    from tests.host_emul import emul_lib
    return emul_lib.load()
