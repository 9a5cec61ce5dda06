import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def _tuplify(spec):
    """json turns tuples into lists; rebuild the oracle-style spec."""
    kind = spec[0]
    if kind in ("union", "intersection", "difference"):
        return (kind, [_tuplify(c) for c in spec[1]], spec[2])
    if kind == "repeat":
        return (kind, tuple(spec[1]), _tuplify(spec[2]), list(spec[3]))
    prm = dict(spec[1])
    if "bbox" in prm:
        prm["bbox"] = tuple(prm["bbox"])
    return (kind, prm)


def load_sdf_specs():
    with open(os.path.join(GOLDEN, "sdf_cases.json")) as f:
        return [_tuplify(s) for s in json.load(f)]


@pytest.fixture(scope="session")
def sdf_specs():
    return load_sdf_specs()


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(1.0, float(np.max(np.abs(b)))) if b.size else 1.0
    return float(np.max(np.abs(a - b)) / scale) if a.size else 0.0


# --------------------------------------------------------------------------------------------
# hostsim: the per-point arithmetic header of the CUDA library compiled for the host (g++), so
# the SDF lowering + interpreter arithmetic can be unit-tested without a GPU.  Test-only.
# --------------------------------------------------------------------------------------------
import ctypes as C  # noqa: E402
import subprocess  # noqa: E402


@pytest.fixture(scope="session")
def hostsim():
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
    out = os.path.join(ROOT, "tests", "hostsim", "libhostsim.so")
    hdr = os.path.join(ROOT, "seismicmesh_b200", "csrc", "dm_sdf.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(
            ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w",
             "-I" + os.path.join(ROOT, "include"), src, "-o", out]
        )
    return C.CDLL(out)


def np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)
