"""Single-GPU SpMV when x no longer fits L2 (the regime of a C4 slice): C2 scaled by `scale` in nodes and links."""
import os, sys, ctypes as C; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs, bench
from recommendersystems_b200 import _native as N
c = rs.widen_float(0.15)
for scale in [float(x) for x in sys.argv[1].split(",")]:
    g = rs.Graph.synthetic(bench.scaled_spec(scale), hub_entries=int(os.environ.get("HUB", "-1"))); g.buildGraph()
    seed = int(np.flatnonzero(g.degrees(raw=True)[:1000] >= 2)[0])
    for prec, pn, vb in ((rs.FP64, "fp64", 8), (rs.FP32, "fp32", 4)):
        for mode in (0, 1):
            os.environ["RWR_DEBUG_MODE"] = str(mode)
            a, b = C.c_float(), C.c_float()
            assert N.lib().rwr_profile_iteration(g._h, seed, c, prec, 10, C.byref(a), C.byref(b)) == 0
            i = g.info()
            alg = i.nnz * (4 + vb) + 4 * (i.n_nodes + 1) + 2 * i.n_nodes * vb
            print(f"scale {scale} n {i.n_nodes} x {i.n_nodes * vb / 1e6:.0f} MB nnz {i.nnz} {pn} mode {mode}: spmv {a.value:.4f} ms finish {b.value:.4f} ms -> "
                  f"{i.nnz / a.value / 1e6:.1f} G links/s, {alg / a.value / 1e6 / 6515.7:.3f} of roofline", flush=True)
    g.close()
