import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The native libraries are built in-tree by __graft_entry__.build() and are not in the git history; a run in a fresh
    # checkout builds what is missing first (nvcc cross-compiles without a GPU; two minutes).  A stale library is the
    # build script's business (codecad_b200/build.py compares time stamps), not this hook's.
    lib = os.path.join(ROOT, "codecad_b200", "libcodecad_b200.so")
    oracle_lib = os.path.join(ROOT, "oracle", "liboracle.so")
    if not (os.path.exists(lib) and os.path.exists(oracle_lib)):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def scenes():
    from scenes import load_scenes
    return load_scenes()
