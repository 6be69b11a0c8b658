import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
DIAG = os.path.join(ROOT, "gpurun_out", "diag")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_cases():
    with open(os.path.join(GOLDEN, "index.json")) as f:
        return [c["name"] for c in json.load(f)["cases"]]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    with open(os.path.join(GOLDEN, f"{name}.json")) as f:
        lists = json.load(f)
    return {k: z[k] for k in z.files}, lists


def dump_diag(tag, **arrays):
    """Keep mismatching arrays for offline inspection (gpurun_out/ travels back from the GPU box)."""
    try:
        os.makedirs(DIAG, exist_ok=True)
        np.savez_compressed(os.path.join(DIAG, f"{tag}.npz"), **arrays)
    except Exception:
        pass


def assert_same(got, want, what, tag=None):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    if np.array_equal(got, want):
        return
    bad = np.argwhere(got != want)
    if tag:
        dump_diag(tag, got=got, want=want)
    d = np.abs(got.astype(np.int64) - want.astype(np.int64)) if got.dtype.kind in "iub" else np.abs(got - want)
    raise AssertionError(f"{what}: {len(bad)} of {got.size} elements differ (max |diff| {d.max()}); "
                         f"first at {bad[:5].tolist()}: got {[got[tuple(i)].item() for i in bad[:5]]} "
                         f"want {[want[tuple(i)].item() for i in bad[:5]]}")


def angle_diff(a, b):
    """difference on the circle of period pi"""
    d = np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)) % np.pi
    return np.minimum(d, np.pi - d)


@pytest.fixture(scope="session")
def hostcheck():
    """g++ build of the FPB_HD device routines (test infrastructure only)."""
    import ctypes
    import subprocess
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "hostcheck")], check=True, capture_output=True)
    lib = ctypes.CDLL(os.path.join(ROOT, "tests", "_build", "libhostcheck.so"))
    lib.hc_patch_otsu.restype = ctypes.c_float
    return lib
