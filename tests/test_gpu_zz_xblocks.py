"""Experimental column blocking of x (RWR_X_BLOCKS): written when no GPU time was left in round 1,
so it is off by default and this test may fail without failing the suite.  It runs last (file name) and in a process of
its own (tests/xblocks_worker.py), with a time limit, so that nothing it does can disturb the other GPU tests."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.xfail(strict=False, reason="experimental path, not yet validated on a GPU (off by default)")
@pytest.mark.parametrize("blocks", [4, 7])
def test_column_blocking_matches_the_oracle(blocks):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "xblocks_worker.py"), str(blocks)], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-2000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0 and "xblocks ok" in r.stdout
