"""One SpMM tile (8 FP64 / 16 FP32 seeds) of 20 iterations on the C2 graph: ncu target for k_spmm."""
import sys; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs, bench
prec = rs.FP32 if (len(sys.argv) > 1 and sys.argv[1] == "fp32") else rs.FP64
g = rs.Graph.synthetic(bench.C2_SPEC); g.buildGraph()
seeds = bench.pick_seeds(g.degrees(raw=True), bench.C2_SPEC["n_users"], 16 if prec == rs.FP32 else 8)
rec = rs.Recommender(g, prec)
rec.RecommendationBatch(seeds, 0.15, 20, 10)
print("iterate_ms", rec.last_info.iterate_ms)
