"""Worker of tests/test_gpu_zz_xblocks.py: runs in its own process with RWR_X_BLOCKS set, so that the experimental
column blocking of x cannot disturb the CUDA context of the main test process.
Usage: python tests/xblocks_worker.py <blocks>      exit code 0 = every check passed."""
import os
import sys

BLOCKS = int(sys.argv[1]) if len(sys.argv) > 1 else 4
os.environ["RWR_X_BLOCKS"] = str(BLOCKS)

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from conftest import C1_SPEC  # noqa: E402
import oracle as O  # noqa: E402
import recommendersystems_b200 as rs  # noqa: E402
from recommendersystems_b200 import _native as N  # noqa: E402
from test_gpu_parity import assert_close_fp32, assert_close_fp64, same_ranking  # noqa: E402

C015 = rs.widen_float(0.15)


def check(spec, opts):
    cpu = O.synth_generate(spec)
    og = O.OracleGraph(cpu["node_id"], cpu["node_type"], cpu["src"], cpu["dst"], cpu["etype"], cpu["w"])
    assert og.build() == 0
    gg = rs.Graph.synthetic(spec, **opts)
    gg.buildGraph()
    info = gg.info()
    assert info.x_blocks > 1, f"column blocking not active (x_blocks = {info.x_blocks})"
    deg = np.bincount(cpu["src"], minlength=og.n)
    seeds = [int(s) for s in np.flatnonzero((deg > 0) & (np.arange(og.n) < spec["n_users"]))[:2]] + [int(np.argmax(deg))]
    for seed in seeds:
        for it in (1, 2, 20):
            want, _ = og.run(seed, C015, n_iter=it)
            m = rs.Model(gg, C015, seed); m.run(it)
            assert_close_fp64(m.rank, want, f"blocked seed {seed} iter {it} {opts}")
        want, _ = og.run(seed, C015, n_iter=20)
        m32 = rs.Model(gg, C015, seed, precision=rs.FP32); m32.run(20)
        assert_close_fp32(m32.rank, want, f"blocked fp32 seed {seed}")
        ids, sc = og.recommend(seed, 0.15, 20, top_n=10)
        top = rs.Recommender(gg).Recommendation(seed, 0.15, 20, 10)
        same_ranking([p[0] for p in top], [p[1] for p in top], ids, sc)
        thr = 1e-6 * og.n
        _, want_it = og.run(seed, C015, threshold=thr)
        m = rs.Model(gg, C015, seed); m.run(thr)
        assert m.nIterations == want_it, (seed, thr, m.nIterations, want_it)
    return info.x_blocks


def main():
    got = []
    got.append(check(C1_SPEC, dict()))                                              # valued layout (fractional weights)
    got.append(check(dict(C1_SPEC, n_mention=0, seed=77), dict()))                  # index-only layout
    got.append(check(dict(C1_SPEC, n_mention=0, seed=78), dict(relabel=False, hub_entries=0)))
    big = dict(C1_SPEC, seed=5, n_users=40_000, n_items=360_000, n_third=0, n_like=1_400_000, n_friend=300_000, n_follow=0,
               n_mention=0, undefined_per_mille=0)                                 # ~3.4 M links: many tiles per block
    got.append(check(big, dict()))
    print("xblocks ok", BLOCKS, got)


if __name__ == "__main__":
    main()
