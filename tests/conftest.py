"""pytest configuration: `gpu` marker, import path, shared helpers.

-m "not gpu": oracle vs golden vectors, host logic, C-ABI load/exports (no compute calls without a GPU).
-m gpu      : parity of the CUDA path (through the C-ABI) against the CPU oracle.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    O.lib()
    return O


@pytest.fixture(scope="session")
def model_dir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("models"))


def rand_blocks(qtype: int, n_blocks: int, rng: np.random.Generator, wild: bool = False) -> np.ndarray:
    """Random packed blocks [n_blocks, block_bytes].  wild=True leaves the f16 scale fields fully random
    (NaN/Inf/subnormal included); otherwise they are set to small finite values."""
    from oracle import oracle as O
    _, bb = O.BLOCK[qtype]
    raw = rng.integers(0, 256, size=(n_blocks, bb), dtype=np.uint8)
    if wild or n_blocks == 0:
        return raw

    def setf16(col, vals):
        raw[:, col:col + 2] = vals.astype(np.float16).view(np.uint8).reshape(-1, 2)

    if qtype in (O.Q8_0, O.Q4_0, O.Q5_0):
        setf16(0, rng.normal(0, 0.01, n_blocks))
    elif qtype in (O.Q4_K, O.Q5_K):
        setf16(0, np.abs(rng.normal(0, 0.01, n_blocks)))
        setf16(2, np.abs(rng.normal(0, 0.05, n_blocks)))
    elif qtype == O.Q6_K:
        setf16(208, rng.normal(0, 0.01, n_blocks))
    return raw
