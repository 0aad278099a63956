import os
import sys

import pytest

# Ranks-as-threads tests (tests/test_gpu_comm.py) co-run kernels of several contexts on one GPU: a put kernel of one rank
# waits for a flag another rank's kernel raises.  With CUDA's default LAZY module loading the first launch of a kernel
# synchronises the context, i.e. it would wait for that spinning put kernel (the case the CUDA programming guide warns
# about under "Lazy Loading / concurrent execution").  Load everything up front.  One process per GPU is not affected.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests only make sense where a GPU exists; there is no fallback to run instead
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (run with gpurun)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from tests.oracle_lib import load_oracle
    return load_oracle()
