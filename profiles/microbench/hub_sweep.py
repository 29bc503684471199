"""SpMV time vs the size of the shared-memory hub table (the rest of the 256 KB per SM stays L1: outstanding-miss capacity)."""
import os, sys, ctypes as C; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs, bench
from recommendersystems_b200 import _native as N
c = rs.widen_float(0.15)
g = None
for hub in [int(x) for x in sys.argv[1].split(",")]:
    g = rs.Graph.synthetic(bench.C2_SPEC, hub_entries=hub); g.buildGraph()
    seed = int(np.flatnonzero(g.degrees(raw=True)[:1000] >= 2)[0])
    out = []
    for prec, pn in ((rs.FP64, "fp64"), (rs.FP32, "fp32")):
        for mode in (0, 1):
            os.environ["RWR_DEBUG_MODE"] = str(mode)
            a, b = C.c_float(), C.c_float()
            assert N.lib().rwr_profile_iteration(g._h, seed, c, prec, 10, C.byref(a), C.byref(b)) == 0
            out.append(f"{pn} m{mode} {a.value:.4f}")
    i = g.info()
    print(f"hub {hub} ({i.hub_entries_fp64}/{i.hub_entries_fp32}): " + "  ".join(out), flush=True)
    g.close()
