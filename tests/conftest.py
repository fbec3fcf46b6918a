import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def fnv1a_u32(a: np.ndarray) -> str:
    """FNV-1a over the little-endian bytes of a uint32 array (same as tests/golden/make_golden.cpp)."""
    h = 1469598103934665603
    data = np.ascontiguousarray(a, dtype=np.uint32).view(np.uint8)
    # chunked pure-python would be slow for 1M+; use the recurrence in numpy-free C-like loop via int ops on bytes
    mask = (1 << 64) - 1
    for b in data.tobytes():
        h = ((h ^ b) * 1099511628211) & mask
    return f"{h:016x}"


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    import oracle as _oracle

    _oracle.build()
    return _oracle


@pytest.fixture(scope="session")
def glu():
    """The product package (ctypes over gl-radix-sort_b200/libglu_b200.so). Raises if the library is missing."""
    import __graft_entry__ as entry

    return entry.load_package()


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu and needs a CUDA device; no CPU fallback exists")
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)


def to_device(a: np.ndarray, device):
    """Upload a numpy array (uint32 goes through an int32 view: torch has no native uint32 arithmetic)."""
    import torch

    if a.dtype == np.uint32:
        return torch.from_numpy(a.view(np.int32).copy()).to(device)
    return torch.from_numpy(a.copy()).to(device)


def to_host(t, dtype) -> np.ndarray:
    a = t.detach().cpu().numpy()
    return a.view(dtype) if a.dtype != dtype else a
