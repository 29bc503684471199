"""SpMV time vs the footprint of the gathered vector: uniform-degree synthetic graphs, same link count, fewer nodes."""
import os, sys, ctypes as C; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs, bench
from recommendersystems_b200 import _native as N
c = rs.widen_float(0.15)
p1 = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for frac in (0.1, 0.2, 0.35, 0.5, 0.65, 0.8, 1.0):
    spec = dict(bench.C2_SPEC); spec["p1_byte"] = p1
    spec["n_users"] = int(1_000_000 * frac); spec["n_items"] = int(10_000_000 * frac)
    g = rs.Graph.synthetic(spec, hub_entries=0 if p1 == 128 else -1); g.buildGraph()
    seed = int(np.flatnonzero(g.degrees(raw=True)[:1000] >= 2)[0])
    for prec, pn in ((rs.FP64, "fp64"), (rs.FP32, "fp32")):
        os.environ["RWR_DEBUG_MODE"] = "1"
        a, b = C.c_float(), C.c_float()
        assert N.lib().rwr_profile_iteration(g._h, seed, c, prec, 10, C.byref(a), C.byref(b)) == 0
        i = g.info()
        mb = i.n_nodes * (8 if prec == rs.FP64 else 4) / 1e6
        print(f"p1 {p1} n {i.n_nodes} x {mb:.0f} MB {pn}: nnz {i.nnz} spmv(mode 1) {a.value:.4f} ms -> {i.nnz / a.value / 1e6:.1f} G gathers/s", flush=True)
    g.close()
