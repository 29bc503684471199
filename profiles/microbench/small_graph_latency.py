"""Request latency on the reference's own scale (C1: an ego network of ~10 k nodes / ~100 k links): Recommendation(seed,
0.15f, 20, 10) and Model.run(20), with the iteration loop replayed as a CUDA graph (default) or launched kernel by
kernel (RWR_NO_GRAPH=1)."""
import os, sys, time; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs
from recommendersystems_b200.rwr import run_fixed
spec = dict(seed=20260101, n_users=1000, n_items=9000, n_third=200, authorship_per_mille=800, n_like=36000, n_friend=8000,
            n_follow=600, n_mention=0, undefined_per_mille=100, scramble=1, p1_byte=61)
g = rs.Graph.synthetic(spec); g.buildGraph()
deg = g.degrees(raw=True)
users = np.flatnonzero(deg[:1000] > 0)
rec = rs.Recommender(g)
c = rs.widen_float(0.15)
for _ in range(20): rec.Recommendation(int(users[0]), 0.15, 20, 10)
t0 = time.perf_counter()
for i in range(200): rec.Recommendation(int(users[i % len(users)]), 0.15, 20, 10)
dt = (time.perf_counter() - t0) / 200
m = run_fixed(g, [int(users[0])], c, 20)
for i in range(20): m.rerun([int(users[i])], c, 20)
t0 = time.perf_counter(); it_ms = 0.0
for i in range(200):
    m.rerun([int(users[i % len(users)])], c, 20); it_ms += m.info().iterate_ms
dt2 = (time.perf_counter() - t0) / 200
print(f"graph {'off' if os.environ.get('RWR_NO_GRAPH') else 'on '}: N={g.info().n_nodes} nnz={g.info().nnz}: Recommendation {dt*1e6:.0f} us per request; "
      f"Model.run(20) {dt2*1e6:.0f} us wall, {it_ms/200*1e3:.0f} us on the device for the 20 iterations")
