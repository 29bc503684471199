import sys; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs, bench
g = rs.Graph.synthetic(bench.C2_SPEC); g.buildGraph()
seeds = bench.pick_seeds(g.degrees(raw=True), bench.C2_SPEC["n_users"], 128)
rec = rs.Recommender(g, rs.FP32)
for s0 in range(0, 128, 16):
    rec.RecommendationBatch(seeds[s0:s0 + 16], 0.15, 20, 10)
    print("fp32 tile", s0 // 16, round(rec.last_info.iterate_ms, 1), "ms")
rec.RecommendationBatch(seeds, 0.15, 20, 10)
print("fp32 all 128:", round(rec.last_info.iterate_ms, 1), "ms")
rec = rs.Recommender(g, rs.FP64)
rec.RecommendationBatch(seeds, 0.15, 20, 10)
print("fp64 all 128:", round(rec.last_info.iterate_ms, 1), "ms")
