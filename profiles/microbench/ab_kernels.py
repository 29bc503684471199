"""A/B of the SpMV kernel variants on the C2 graph: kernel 0 (warp-streamed), 1 (pipelined), 2 (phased); FP64 and FP32."""
import sys, ctypes as C; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs, bench
from recommendersystems_b200 import _native as N
c = rs.widen_float(0.15)
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
kerns = [int(k) for k in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 2]
for kern in kerns:
    g = rs.Graph.synthetic(bench.scaled_spec(scale), kernel=kern); g.buildGraph()
    seed = int(np.flatnonzero(g.degrees(raw=True)[:1000] >= 2)[0])
    for prec, name in ((rs.FP64, "fp64"), (rs.FP32, "fp32")):
        a, b = C.c_float(), C.c_float()
        assert N.lib().rwr_profile_iteration(g._h, seed, c, prec, 20, C.byref(a), C.byref(b)) == 0, N.last_error()
        i = g.info()
        print(f"kernel {kern} {name}: spmv {a.value:.4f} ms fixup {b.value:.4f} ms -> {i.nnz / a.value / 1e6:.1f} GTEPS  "
              f"hub {i.hub_entries_fp64}/{i.hub_entries_fp32} build {i.build_ms:.0f} ms", flush=True)
    g.close()
