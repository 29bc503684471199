"""C5 (BASELINE configs[4]) in one script: Experiment-style evaluation on a 10 M-node synthetic graph -- hold out the
newest 10 % of the test users' likes, top-10 for every test user through the batched path, recall@10 and seeds/s.
usage: python profiles/microbench/eval_c5.py [n_test_users=2048] [scale=1.0]   (the full configuration is 100000 users)"""
import sys, time; sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import numpy as np, recommendersystems_b200 as rs
from recommendersystems_b200.experiment import hold_out_likes, recall_at_k
n_users = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
spec = dict(seed=20260105, n_users=int(1_000_000 * scale), n_items=int(9_000_000 * scale), n_third=0, authorship_per_mille=1000,
            n_like=int(70_000_000 * scale), n_friend=int(20_000_000 * scale), n_follow=0, n_mention=0, undefined_per_mille=0,
            scramble=1, p1_byte=61)
t0 = time.perf_counter()
g0 = rs.Graph.synthetic(spec); links = g0.export_links(); g0.close()
like_deg = np.bincount(links["src"][links["etype"] == 1], minlength=len(links["node_id"]))
cand = np.flatnonzero(like_deg[:spec["n_users"]] >= 20)
users = cand[(np.arange(n_users) * 7919) % len(cand)]
users = np.unique(users)
held, test = hold_out_likes(links, users, 0.1)
g = rs.Graph.from_arrays(held["node_id"], held["node_type"], held["src"], held["dst"], held["etype"], held["w"]); g.buildGraph()
print(f"graph {g.info().n_nodes} nodes {g.info().nnz} links, {len(users)} test users, {sum(len(t) for t in test.values())} held-out likes; "
      f"setup {time.perf_counter() - t0:.1f} s", flush=True)
for prec, pn in ((rs.FP64, "fp64"), (rs.FP32, "fp32")):
    rec = rs.Recommender(g, prec)
    rec.RecommendationBatch(users[:16], 0.15, 20, 10)
    t0 = time.perf_counter()
    ids, sc, cnt = rec.RecommendationBatch(users, 0.15, 20, 10)
    dt = time.perf_counter() - t0
    r, hits, counted = recall_at_k(ids, cnt, users, test)
    print(f"{pn}: {len(users)} users in {dt:.2f} s -> {len(users) / dt:.1f} seeds/s; recall@10 {r:.4f}, hits {hits}, users counted {counted}", flush=True)
