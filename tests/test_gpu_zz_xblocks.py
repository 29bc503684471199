"""Experimental column blocking of x (RWR_X_BLOCKS), off by default.  Written at the very end of
round 1: one GPU run with 4 blocks passed (profiles/microbench/xblocks_parity_r01.log); other block counts have not run
yet and may fail without failing the suite.  The check runs last (file name) and in a process of its own
(tests/xblocks_worker.py: the knob is read from the environment when a graph is built), with a time limit."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("blocks", [4, pytest.param(7, marks=pytest.mark.xfail(
    strict=False, reason="experimental path: this block count has not run on a GPU yet"))])
def test_column_blocking_matches_the_oracle(blocks):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "xblocks_worker.py"), str(blocks)], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-2000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0 and "xblocks ok" in r.stdout
