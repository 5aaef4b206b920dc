import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with `-m gpu` on the GPU box)')


@pytest.fixture(scope='session')
def oracle():
    from oracle import oracle as orc
    orc.lib()
    return orc


@pytest.fixture(scope='session')
def gpu_ctx():
    """One context for the whole GPU session; creating it fails loudly when the CUDA library or the GPU is missing."""
    import nlos_surface_optimization_b200 as nb
    return nb.default_context(0)
