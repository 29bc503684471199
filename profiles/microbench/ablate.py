"""Ablations of the phased SpMV kernel (measurement only): which part of the work costs what."""
import os, sys, ctypes as C; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs, bench
from recommendersystems_b200 import _native as N
c = rs.widen_float(0.15)
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
g = rs.Graph.synthetic(bench.scaled_spec(scale), kernel=1, hot_min_degree=2); g.buildGraph()
seed = int(np.flatnonzero(g.degrees(raw=True)[:1000] >= 2)[0])
names = {0: "full kernel", 1: "no gathers (stream + smem + rows)", 2: "all gathers from shared hub", 3: "all gathers global, 32 KB window",
         5: "all gathers global, 32 MB window", 4: "full gathers, no row phase"}
for mode, nm in names.items():
    os.environ["RWR_DEBUG_MODE"] = str(mode)
    a, b = C.c_float(), C.c_float()
    assert N.lib().rwr_profile_iteration(g._h, seed, c, rs.FP64, 10, C.byref(a), C.byref(b)) == 0
    print(f"mode {mode} {nm:36s}: spmv {a.value:.4f} ms -> {g.info().nnz / a.value / 1e6:.1f} GTEPS")
