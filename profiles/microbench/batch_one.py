import sys; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs, bench
prec = rs.FP32 if sys.argv[1] == "fp32" else rs.FP64
g = rs.Graph.synthetic(bench.C2_SPEC); g.buildGraph()
seeds = bench.pick_seeds(g.degrees(raw=True), bench.C2_SPEC["n_users"], 16)
rec = rs.Recommender(g, prec)
rec.RecommendationBatch(seeds, 0.15, 4, 10)
print(rec.last_info.iterate_ms)
