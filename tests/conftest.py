import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name: str):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


@pytest.fixture(scope="session")
def built_library():
    """The C-ABI library, compiled in-tree if missing or stale (nvcc cross-compiles without a GPU)."""
    from multi_stylegan_b200 import _lib
    _lib.build()
    return _lib.lib()


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max-abs-err / max-abs-ref (SURVEY.md §8d parity metric)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = b.abs().max().item()
    return (a - b).abs().max().item() / (denom if denom > 0 else 1.0)


@pytest.fixture(autouse=True)
def _restore_global_torch_flags():
    """ModelWrapper switches torch.backends.cuda.matmul.allow_tf32 on for the process (the reference environment's
    default); tests that compare library fp32 matmuls at 1e-5 must not inherit it from an earlier test."""
    saved = torch.backends.cuda.matmul.allow_tf32
    yield
    torch.backends.cuda.matmul.allow_tf32 = saved


@pytest.fixture()
def oracle_backend(monkeypatch):
    """Swap the device ops for the CPU oracle so the *host logic* (autograd wiring, module tree,
    train step, data-parallel plumbing) can be exercised without a GPU.  Test-only."""
    from tests import backend_oracle
    from multi_stylegan_b200 import _C
    for name in backend_oracle.__all__:
        monkeypatch.setattr(_C, name, getattr(backend_oracle, name))
    yield
