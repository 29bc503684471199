import sys; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs, bench
g = rs.Graph.synthetic(bench.C2_SPEC); g.buildGraph()
seeds = bench.pick_seeds(g.degrees(raw=True), bench.C2_SPEC["n_users"], 16)
for prec, name in ((rs.FP32, "fp32"), (rs.FP64, "fp64")):
    rec = rs.Recommender(g, prec)
    for it in (2, 4, 8, 12, 16, 20):
        rec.RecommendationBatch(seeds, 0.15, it, 10)
        print(name, it, "iterations:", round(rec.last_info.iterate_ms, 2), "ms ->", round(rec.last_info.iterate_ms / it, 2), "ms/iter")
