import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def golden_json(arr):
    return json.loads(bytes(arr.tolist()).decode())


@pytest.fixture(scope="session")
def host_harness():
    """tests/host_harness.cpp compiled for the CPU: the product's box_geom.cuh geometry without a GPU."""
    import ctypes
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhostharness.so")
    src = os.path.join(ROOT, "tests", "host_harness.cpp")
    hdrs = [os.path.join(ROOT, "video_text_detection_system_b200", "csrc", h) for h in ("box_geom.cuh", "resize_tab.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(p) for p in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-x", "c++", src,
                               "-o", so])
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def lib_built():
    from video_text_detection_system_b200.build import build_library
    return build_library()
