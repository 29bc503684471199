"""Column blocking of x with padded virtual rows (rwr_opts.x_blocks / RWR_X_BLOCKS), off by default: parity with 4 and 7
blocks on four graphs.  Runs in a process of its own (tests/xblocks_worker.py: the knob is read from the environment when
a graph is built), with a time limit."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("blocks", [4, 7])
def test_column_blocking_matches_the_oracle(blocks):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "xblocks_worker.py"), str(blocks)], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-2000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0 and "xblocks ok" in r.stdout
