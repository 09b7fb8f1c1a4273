import ctypes
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "node_golden.npz"))
    meta = json.loads(bytes(z["meta_json"]).decode())
    return z, meta


@pytest.fixture(scope="session")
def golden_cases():
    from oracle.gen_golden import golden_cases as gc
    return gc()


@pytest.fixture(scope="session")
def host_harness():
    """g++ build of the __host__ __device__ node primitives (no CUDA needed)."""
    out = os.path.join(ROOT, "build", "libhost_harness.so")
    src = os.path.join(ROOT, "tests", "host_harness.cpp")
    hdr = os.path.join(ROOT, "circuitvision_b200", "csrc", "node_prims.cuh")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if (not os.path.exists(out)) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src])
    return ctypes.CDLL(out)


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def terminal_golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "terminal_golden.npz"))
    return json.loads(bytes(z["meta_json"]).decode())


@pytest.fixture(scope="session")
def terminal_cases():
    from oracle.gen_golden import terminal_cases as tc
    return tc()
