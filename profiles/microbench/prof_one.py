"""One profile_iteration call of a chosen kernel variant on the C2 graph (ncu target)."""
import sys, ctypes as C; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs, bench
from recommendersystems_b200 import _native as N
kern = int(sys.argv[1]); hmin = int(sys.argv[2]); prec = rs.FP32 if (len(sys.argv) > 3 and sys.argv[3] == "fp32") else rs.FP64
g = rs.Graph.synthetic(bench.C2_SPEC, hot_min_degree=hmin); g.buildGraph()
seed = int(np.flatnonzero(g.degrees(raw=True)[:1000] >= 2)[0])
a, b = C.c_float(), C.c_float()
assert N.lib().rwr_profile_iteration(g._h, seed, rs.widen_float(0.15), prec, 5, C.byref(a), C.byref(b)) == 0
print("spmv ms", a.value, "fixup ms", b.value)
