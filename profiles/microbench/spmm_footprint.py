"""Batched path: time per link of one SpMM pass vs the size of X[n, B] (same link count, fewer nodes, uniform degrees or
the C2 power law): is the pass bound by HBM random accesses or by the kernel?"""
import sys; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs, bench
for p1 in (61, 128):
    for frac in (0.05, 0.1, 0.2, 0.5, 1.0):
        spec = dict(bench.C2_SPEC); spec["p1_byte"] = p1
        spec["n_users"] = int(1_000_000 * frac); spec["n_items"] = int(10_000_000 * frac)
        g = rs.Graph.synthetic(spec); g.buildGraph()
        seeds = bench.pick_seeds(g.degrees(raw=True), spec["n_users"], 16)
        rec = rs.Recommender(g)
        best = 1e9
        for _ in range(3):
            rec.RecommendationBatch(seeds, 0.15, 20, 10)
            best = min(best, rec.last_info.iterate_ms / 2 / 20)
        i = g.info()
        print(f"p1 {p1} n {i.n_nodes} X {i.n_nodes * 64 / 1e6:.0f} MB nnz {i.nnz}: {best:.4f} ms per pass -> {i.nnz / best / 1e6:.1f} G links/s", flush=True)
        g.close()
