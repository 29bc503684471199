"""SpMV time of the valued layout (4 + vb bytes per link) against the index-only layout on the C2 graph."""
import sys, ctypes as C; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs, bench
from recommendersystems_b200 import _native as N
c = rs.widen_float(0.15)
for layout, nm in ((N.LAYOUT_INDEX, "index-only"), (N.LAYOUT_VALUED, "valued")):
    g = rs.Graph.synthetic(bench.C2_SPEC, layout=layout); g.buildGraph()
    seed = int(np.flatnonzero(g.degrees(raw=True)[:1000] >= 2)[0])
    for prec, pn, vb in ((rs.FP64, "fp64", 8), (rs.FP32, "fp32", 4)):
        a, b = C.c_float(), C.c_float()
        assert N.lib().rwr_profile_iteration(g._h, seed, c, prec, 20, C.byref(a), C.byref(b)) == 0
        i = g.info()
        alg = i.nnz * (4 + vb) + 4 * (i.n_nodes + 1) + 2 * i.n_nodes * vb
        print(f"{nm} {pn}: spmv {a.value:.4f} ms epilogue {b.value:.4f} ms -> {alg / a.value / 1e6:.0f} GB/s algorithmic ({alg / a.value / 1e6 / 6515.7:.3f} of 6515.7), "
              f"{i.nnz / (a.value + b.value) / 1e6:.1f} GTEPS per iteration", flush=True)
    g.close()
