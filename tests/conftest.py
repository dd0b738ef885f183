import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def b2pt():
    """The product binding; builds libb2pt.so if it is missing (nvcc cross-compiles without a GPU)."""
    from raytracingtherestofyourlife_b200 import _build
    _build.build_lib()
    import raytracingtherestofyourlife_b200 as B
    B.lib()
    return B


@pytest.fixture(scope="session")
def gpu_ctx(b2pt):
    """One context on cuda:0 with the Cornell scene; fails loudly (no fallback) when no B200 is present."""
    ctx = b2pt.Context(0)
    ctx.set_scene(b2pt.Scene.cornell())
    ctx.build_bvh()
    yield ctx
    ctx.close()
