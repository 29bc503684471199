"""C3: batched RWR for many seeds as SpMM tiles on the C2 graph -> seeds/s and seed-GTEPS."""
import sys, time; sys.path.insert(0, ".")
import numpy as np, recommendersystems_b200 as rs, bench
n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = rs.Graph.synthetic(bench.C2_SPEC); g.buildGraph()
deg = g.degrees(raw=True)
seeds = bench.pick_seeds(deg, bench.C2_SPEC["n_users"], n_seeds)
for prec, name in ((rs.FP64, "fp64"), (rs.FP32, "fp32")):
    rec = rs.Recommender(g, prec)
    rec.RecommendationBatch(seeds[:16], 0.15, 20, 10)
    t0 = time.perf_counter()
    ids, sc, cnt = rec.RecommendationBatch(seeds, 0.15, 20, 10)
    dt = time.perf_counter() - t0
    i = rec.last_info
    print(f"{name}: {n_seeds} seeds in {dt:.3f} s -> {n_seeds/dt:.1f} seeds/s; iterate {i.iterate_ms:.1f} ms total {i.total_ms:.1f} ms "
          f"-> {g.info().nnz*20*n_seeds/i.iterate_ms/1e6:.1f} seed-GTEPS (loop) {g.info().nnz*20*n_seeds/dt/1e9:.1f} e2e; launches {i.kernel_launches}")
