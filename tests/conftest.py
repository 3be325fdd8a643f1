"""Shared test plumbing: markers, import paths, golden loaders.

`-m "not gpu"` covers the oracle against the golden vectors, the host logic and the C-ABI
export table; `-m gpu` holds the parity tests proper (CUDA path vs oracle, through the C-ABI).
"""

import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "fast-forward-indexes_b200")
ORACLE = os.path.join(ROOT, "oracle")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (PKG, ORACLE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_kat():
    with open(os.path.join(GOLDEN, "ref_kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_random():
    with open(os.path.join(GOLDEN, "ref_random.json")) as f:
        meta = json.load(f)
    arrays = np.load(os.path.join(GOLDEN, "ref_random.npz"))
    return meta, arrays


@pytest.fixture(scope="session")
def golden_es():
    with open(os.path.join(GOLDEN, "ref_es.json")) as f:
        meta = json.load(f)
    arrays = np.load(os.path.join(GOLDEN, "ref_es.npz"))
    return meta, arrays


@pytest.fixture(scope="session")
def oracle_c():
    """ctypes handle on the plain-C oracle (built on demand; test infrastructure)."""
    import ctypes
    import subprocess

    so = os.path.join(ORACLE, "libff_oracle.so")
    src = os.path.join(ORACLE, "ff_oracle_c.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE, "-s"])
    lib = ctypes.CDLL(so)
    lib.ffo_dot_f32.restype = ctypes.c_float
    return lib
