"""C5 (BASELINE configs[4]) in one script: Experiment-style evaluation on a 10 M-node synthetic graph -- device-side hold-out
of the newest tenth of the test users' likes, full-ranking hits / average precision and recall@10 for every test user.
usage: python profiles/microbench/eval_c5.py [n_test_users=2048] [scale=1.0]   (the full configuration is 100000 users)"""
import sys, time; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs, bench
from recommendersystems_b200.experiment import summarize
n_users = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
spec = dict(bench.C5_SPEC)
for k in ("n_users", "n_items", "n_like", "n_friend"):
    spec[k] = int(spec[k] * scale)
t0 = time.perf_counter()
g = rs.Graph.synthetic(spec)
cand = np.flatnonzero(g.degrees(raw=True)[:spec["n_users"]] >= 20)
users = cand[np.unique(np.linspace(0, len(cand) - 1, min(n_users, len(cand))).astype(np.int64))].astype(np.int32)
test = g.hold_out(users, 10, 9)
g.buildGraph()
print(f"graph {g.info().n_nodes} nodes {g.info().nnz} links, {len(users)} test users, {int(g.test_ptr[-1])} held-out likes; "
      f"setup {time.perf_counter() - t0:.1f} s", flush=True)
for prec, pn in ((rs.FP64, "fp64"), (rs.FP32, "fp32")):
    rs.evaluate_users(g, users[:16], None, 0.15, 20, k=10, precision=prec)
    t0 = time.perf_counter()
    r = rs.evaluate_users(g, users, None, 0.15, 20, k=10, precision=prec)
    dt = time.perf_counter() - t0
    print(pn, f"{len(users) / dt:.1f} users/s", summarize(r), flush=True)
