"""Per-phase cycle counts of the warp-streamed SpMV (needs a -DRWR_PROFILE_CLOCKS build of librwr_b200.so)."""
import sys, ctypes as C; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs, bench
from recommendersystems_b200 import _native as N
from recommendersystems_b200.rwr import run_fixed
L = N.lib(); L.rwr_debug_clocks_ws.argtypes = [C.c_void_p, C.c_int]
g = rs.Graph.synthetic(bench.C2_SPEC); g.buildGraph()
seed = int(np.flatnonzero(g.degrees(raw=True)[:1000000] >= 8)[0])
c = rs.widen_float(0.15)
run_fixed(g, [seed], c, 5).close()
L.rwr_debug_clocks_ws(None, 1)
r = run_fixed(g, [seed], c, 10); print("iterate_ms", r.info().iterate_ms / 10); r.close()
out = (C.c_ulonglong * 16)(); L.rwr_debug_clocks_ws(out, 0)
names = {0: "wait idx of next stage", 1: "issue gathers", 2: "issue idx loads", 3: "wait gathers of this stage", 4: "rounds (scan + epilogue)", 5: "tile end"}
stages = (g.info().nnz + g.info().n_nodes) / 256 * 10
for k, nm in names.items(): print(f"{nm:34s} {out[k]/stages:9.0f} cycles/stage/warp")
