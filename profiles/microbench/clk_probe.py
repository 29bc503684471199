import sys, ctypes as C; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs, bench
from recommendersystems_b200 import _native as N
from recommendersystems_b200.rwr import run_fixed
L = N.lib(); L.rwr_debug_clocks.argtypes = [C.c_void_p, C.c_int]
g = rs.Graph.synthetic(bench.C2_SPEC); g.buildGraph()
seed = int(np.flatnonzero(g.degrees(raw=True)[:1000000] >= 8)[0])
c = rs.widen_float(0.15)
run_fixed(g, [seed], c, 5).close()
L.rwr_debug_clocks(None, 1)
r = run_fixed(g, [seed], c, 10); print("iterate_ms", r.info().iterate_ms / 10); r.close()
out = (C.c_ulonglong * 16)(); L.rwr_debug_clocks(out, 0)
names = {0:"P wait prefetched tile(idx)",1:"P issue gathers",2:"P issue next loads (part wait)",3:"P wait empty",4:"P wait gathers+STS",8:"C wait full",9:"C short rows",10:"C barrier1",11:"C long rows",12:"C huge row",13:"C final barrier"}
tiles = g.info().n_chunks * 10
for k, nm in names.items(): print(f"{nm:34s} {out[k]/tiles:9.0f} cycles/tile")
