import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")
    if os.environ.get("CHMY_DRYRUN") == "1":
        # tests/test_gpu_suite_dryrun.py's child process: execute the single-GPU `-m gpu` files on the CPU against
        # tests/dryrun_backend.py (the C ABI restated over the oracle).  Never set on the GPU box.
        from helpers import install_dryrun_if_requested
        install_dryrun_if_requested()


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.lib()
    return o
