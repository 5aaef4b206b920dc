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


@pytest.fixture(scope='session', params=['bvh', 'grid'])
def gpu_ctx(request):
    """One context for the whole GPU session; creating it fails loudly when the CUDA library or the GPU is missing.
    Every GPU test runs twice: with the BVH traversal forward kernel and with the perspective-grid forward kernel (the library's
    `auto` choice picks between them by the number of wall points; the test scenes are small, so both are forced here)."""
    import nlos_surface_optimization_b200 as nb
    ctx = nb.default_context(0)
    ctx.set_option('forward_algo', 1 if request.param == 'bvh' else 2)
    yield ctx
    ctx.set_option('forward_algo', 0)
