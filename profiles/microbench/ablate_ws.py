"""Ablations of the warp-streamed SpMV (measurement only): which part of the work costs what."""
import os, sys, ctypes as C; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs, bench
from recommendersystems_b200 import _native as N
c = rs.widen_float(0.15)
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
g = rs.Graph.synthetic(bench.scaled_spec(scale)); g.buildGraph()
seed = int(np.flatnonzero(g.degrees(raw=True)[:1000] >= 2)[0])
names = {0: "full kernel", 1: "no scan, no stores (every stage summed per lane)", 2: "scan, no y stores", 3: "no gathers (index stream + scan + stores)",
         4: "every gather from the shared hub", 5: "every gather from L2"}
for prec, pn in ((rs.FP64, "fp64"), (rs.FP32, "fp32")):
    for mode, nm in names.items():
        os.environ["RWR_DEBUG_MODE"] = str(mode)
        a, b = C.c_float(), C.c_float()
        assert N.lib().rwr_profile_iteration(g._h, seed, c, prec, 10, C.byref(a), C.byref(b)) == 0
        print(f"{pn} mode {mode} {nm:50s}: spmv {a.value:.4f} ms finish {b.value:.4f} ms -> {g.info().nnz / a.value / 1e6:.1f} GTEPS", flush=True)
