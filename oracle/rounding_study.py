"""TEST INFRASTRUCTURE ONLY -- how much of the FP64 difference between the GPU and the oracle at full C2 size is rounding
noise of the reference's own summation order?

The reference adds the contributions of a row's in-links one by one in source order (Model.cs:85-88); the GPU adds the same
products as a tree.  On a graph whose hubs have 3e5 in-links the sequential sum itself carries ~sqrt(deg) * 2^-53 of
rounding error per iteration.  This script runs one seed for 20 iterations three ways on the full C2 graph -- the oracle
(double, reference order), the oracle's arithmetic in x87 extended precision (a 64-bit-significand yardstick), the GPU
(FP64, single-seed path) -- and prints the pairwise relative differences.   usage: python oracle/rounding_study.py [scale]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import bench
import oracle as O
import recommendersystems_b200 as rs
from recommendersystems_b200.rwr import run_fixed

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
spec = bench.scaled_spec(scale)
g = rs.Graph.synthetic(spec)
g.buildGraph()
links = g.export_links()
og = O.OracleGraph(links["node_id"], links["node_type"], links["src"], links["dst"], links["etype"], links["w"])
raw_deg = np.bincount(links["src"], minlength=og.n)
del links
assert og.build() == 0
seed = int(bench.pick_seeds(raw_deg, spec["n_users"], 1, offset=100_000)[0])     # the first seed of the bench's C3 list
c = O.widen_float(0.15)
t0 = time.perf_counter()
ref, _ = og.run(seed, c, n_iter=20)
t1 = time.perf_counter()
ext = np.empty(og.n, np.float64)
L = O.lib()
L.orc_model_run_extended.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_int32, C.c_void_p]
assert L.orc_model_run_extended(og._h, seed, c, 20, ext.ctypes.data_as(C.c_void_p)) == 0
t2 = time.perf_counter()
m = run_fixed(g, [seed], c, 20, rs.FP64)
gpu = m.scores(0)
m.close()


def rel(a, b, mask):
    return float((np.abs(a[mask] - b[mask]) / np.abs(b[mask])).max())


nz = ext != 0
top = np.argsort(-ext)[:1000]
tmask = np.zeros(og.n, bool); tmask[top] = True
big = ext > 1e-9 * ext.max()
print(f"C2 x {scale}: n {og.n} nnz {og.nnz()} seed {seed}; oracle {t1 - t0:.1f} s, extended {t2 - t1:.1f} s", flush=True)
for name, mask in (("all non-zero scores", nz), ("scores > 1e-9 of the largest", big), ("the 1000 largest scores", tmask)):
    print(f"{name:32s} ({int(mask.sum())} nodes): max rel |oracle - extended| {rel(ref, ext, mask):.3e}   |gpu - extended| {rel(gpu, ext, mask):.3e}   "
          f"|gpu - oracle| {rel(gpu, ref, mask):.3e}", flush=True)
