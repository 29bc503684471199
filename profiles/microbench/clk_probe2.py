import os, sys, ctypes as C; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs, bench
from recommendersystems_b200 import _native as N
L = N.lib(); L.rwr_debug_clocks.argtypes = [C.c_void_p, C.c_int]
g = rs.Graph.synthetic(bench.C2_SPEC, kernel=1, hot_min_degree=2); g.buildGraph()
seed = int(np.flatnonzero(g.degrees(raw=True)[:1000] >= 2)[0])
c = rs.widen_float(0.15)
names = {0:"wait prefetched tile",1:"issue gathers",2:"issue prefetch",3:"single-row tile path",4:"gather wait + STS",5:"barrier 1",6:"short rows",7:"barrier 2",8:"long+huge rows",9:"barrier 3"}
for mode in (0, 1):
    os.environ["RWR_DEBUG_MODE"] = str(mode)
    a, b = C.c_float(), C.c_float()
    L.rwr_profile_iteration(g._h, seed, c, rs.FP64, 3, C.byref(a), C.byref(b))
    L.rwr_debug_clocks(None, 1)
    L.rwr_profile_iteration(g._h, seed, c, rs.FP64, 7, C.byref(a), C.byref(b))
    out = (C.c_ulonglong * 16)(); L.rwr_debug_clocks(out, 0)
    tiles = g.info().n_chunks * 10
    print("mode", mode, "spmv ms", a.value, "total cycles/tile", sum(out[k] for k in names) / tiles)
    for k, nm in names.items(): print(f"   {nm:28s} {out[k]/tiles:9.0f} cycles/tile")
