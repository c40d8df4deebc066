import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mindspore-hp-vae-gan_b200"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")
    # the shared library is a build artefact (git-ignored): on a fresh checkout build it once (nvcc cross-compiles
    # sm_100a without a GPU) instead of failing every test at import time
    if not os.path.exists(os.path.join(ROOT, "mindspore-hp-vae-gan_b200", "hpvg", "libhpvg.so")):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def hpvg_gpu():
    """The product package, initialised on cuda:0.  Fails loudly (no skip, no fallback) when no GPU is visible."""
    import hpvg
    hpvg.init(0)
    return hpvg
