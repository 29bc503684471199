"""One slice of a row-partitioned graph on ONE GPU (RWR_FAKE_COMM: no NCCL, no exchange -- kernel timing only, the
results of the runs are wrong): slice_probe.py <c4|scale> <P> <rank> [fp64|fp32|both] [reps]
Knobs read from the environment by the library: RWR_X_BLOCKS, RWR_PART_NO_HUB, RWR_TILE_LINKS, RWR_DEAL_BLOCK."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, ".")
os.environ["RWR_FAKE_COMM"] = "1"
import numpy as np
import recommendersystems_b200 as rs
import bench
from recommendersystems_b200 import _native as N

what = sys.argv[1] if len(sys.argv) > 1 else "1.0"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
precs = sys.argv[4] if len(sys.argv) > 4 else "fp64"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 20
spec = dict(bench.C4_SPEC) if what == "c4" else bench.scaled_spec(float(what))
comm = rs.Comm(rank, P, b"\0" * 128, device=0) if P > 1 else None
t0 = time.perf_counter()
g = rs.Graph.synthetic(spec, comm=comm)
g.buildGraph()
i = g.info()
c = rs.widen_float(0.15)
out = dict(graph=what, P=P, rank=rank, n=i.n_nodes, nnz=i.nnz, rows=[i.row_begin, i.row_end], x_blocks=i.x_blocks,
           hub64=i.hub_entries_fp64, build_ms=round(i.build_ms, 1), dev_gb=round(i.device_bytes / 1e9, 2),
           setup_s=round(time.perf_counter() - t0, 1), env={k: os.environ.get(k) for k in ("RWR_X_BLOCKS", "RWR_PART_NO_HUB", "RWR_TILE_LINKS")})
for pn, prec in (("fp64", rs.FP64), ("fp32", rs.FP32)):
    if precs not in (pn, "both"):
        continue
    a, b = C.c_float(), C.c_float()
    rc = N.lib().rwr_profile_iteration(g._h, 12345 % i.n_nodes, c, prec, reps, C.byref(a), C.byref(b))
    assert rc == 0, N.last_error()
    vb = 8 if prec == rs.FP64 else 4
    alg = i.nnz / P * (4 + vb) + 4 * (i.n_nodes / P + 1) + i.n_nodes * vb + i.n_nodes / P * vb
    out[pn] = dict(spmv_ms=round(a.value, 4), fin_ms=round(b.value, 4), iter_ms=round(a.value + b.value, 4),
                   hbm_frac=round(alg / ((a.value + b.value) * 1e-3) / 1e9 / bench.measured_peak_gbs()[0], 4))
print(out, flush=True)
g.close()
