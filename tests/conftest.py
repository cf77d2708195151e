import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box via gpurun)")
    config.addinivalue_line("markers", "slow: full-size configs (minutes of host time)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref_check():
    """The reference's own check() when oracle/_ref/shared.so is present (build container, or shipped to the box)."""
    from oracle import load_reference_check
    return load_reference_check()


@pytest.fixture(scope="session")
def golden():
    return json.loads((ROOT / "tests" / "golden" / "check_verdicts.json").read_text())["cases"]


@pytest.fixture(scope="session")
def lib():
    import __graft_entry__ as g
    g.build()
    from mlir_hashjoin_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test on a machine without CUDA: there is no CPU fallback to exercise")
    return torch.device("cuda:0")
