"""SpMV time as a function of the degree skew of the synthetic graph (p1_byte: 61 = R-MAT-like power law, 128 = uniform)."""
import os, sys, ctypes as C; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs, bench
from recommendersystems_b200 import _native as N
c = rs.widen_float(0.15)
for p1 in (128, 96, 61):
    spec = dict(bench.C2_SPEC); spec["p1_byte"] = p1
    g = rs.Graph.synthetic(spec); g.buildGraph()
    seed = int(np.flatnonzero(g.degrees(raw=True)[:1000] >= 2)[0])
    for mode in (0, 1, 3):
        os.environ["RWR_DEBUG_MODE"] = str(mode)
        a, b = C.c_float(), C.c_float()
        assert N.lib().rwr_profile_iteration(g._h, seed, c, rs.FP64, 10, C.byref(a), C.byref(b)) == 0
        i = g.info()
        print(f"p1 {p1} mode {mode}: nnz {i.nnz} max_in {i.max_in_degree} spmv {a.value:.4f} ms finish {b.value:.4f} -> {i.nnz / a.value / 1e6:.1f} GTEPS", flush=True)
    g.close()
