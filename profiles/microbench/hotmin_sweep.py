"""SpMV time vs hot_min_degree (nodes with fewer explicit links are clustered next to their first neighbour)."""
import os, sys, ctypes as C; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs, bench
from recommendersystems_b200 import _native as N
c = rs.widen_float(0.15)
for hm in [int(x) for x in sys.argv[1].split(",")]:
    g = rs.Graph.synthetic(bench.C2_SPEC, hot_min_degree=hm); g.buildGraph()
    seed = int(np.flatnonzero(g.degrees(raw=True)[:1000] >= 2)[0])
    out = []
    for prec, pn in ((rs.FP64, "fp64"), (rs.FP32, "fp32")):
        a, b = C.c_float(), C.c_float()
        assert N.lib().rwr_profile_iteration(g._h, seed, c, prec, 10, C.byref(a), C.byref(b)) == 0
        out.append(f"{pn} spmv {a.value:.4f} finish {b.value:.4f}")
    i = g.info()
    print(f"hot_min {hm} n_hot {i.n_hot} build {i.build_ms:.0f} ms: " + "  ".join(out), flush=True)
    g.close()
