import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """A clean checkout holds no binaries: build librwr_b200.so (nvcc cross-compiles sm_100a without a GPU) and the C++ caller
    before the first test, as `__graft_entry__.build()` does.  The oracle and oracle/_ref build themselves on first use."""
    import subprocess
    pkg = os.path.join(ROOT, "recommendersystems_b200")
    if not os.path.exists(os.path.join(pkg, "librwr_b200.so")):
        subprocess.check_call(["make", "-C", os.path.join(pkg, "csrc"), "-j8", "-s"])
    if not os.path.exists(os.path.join(ROOT, "tests", "cpp", "caller_b200")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s"])


def unhex(xs):
    return np.array([float.fromhex(x) for x in xs], dtype=np.float64)


def load_golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        g = json.load(f)
    if "input" in g:
        g["input"]["w"] = unhex(g["input"]["w"])
    if "graph" in g:
        g["graph"]["w"] = unhex(g["graph"]["w"])
    return g


@pytest.fixture(scope="session")
def golden():
    return load_golden


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


# the C1-shaped synthetic spec used by the CPU and GPU tests
C1_SPEC = dict(seed=20260101, n_users=1000, n_items=9000, n_third=200, authorship_per_mille=800, n_like=36000,
               n_friend=8000, n_follow=600, n_mention=400, undefined_per_mille=100, scramble=1, p1_byte=61, reserved=0)
