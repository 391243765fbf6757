import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    return g.load_package()


@pytest.fixture(scope="session")
def oracle():
    import __graft_entry__ as g
    O = g.load_oracle()
    O.build_c_oracle()
    return O


@pytest.fixture(scope="session")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


GOLDEN = os.path.join(ROOT, "tests", "golden")


def write_report(name, obj):
    """Measured parity figures of a GPU test run, written to gpurun_out/ (scratch on the GPU box; the
    ones that are evidence are copied to profiles/ and committed)."""
    import json
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, name), "w") as f:
            json.dump(obj, f, indent=1, sort_keys=True)
    except OSError:
        pass
