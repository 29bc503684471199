"""A row-partitioned graph that does not fit one GPU (torchrun, one process per GPU): k x C4 -- k * 50 M nodes, k * 1.83 B
links.  No oracle can run at this size; the checks are the invariants of the walk: the ranks sum to N (mass conservation,
Model.cs:44 starts with rank[seed] = N), every rank ends with the same vector, the seed keeps the restart mass.
usage: part_big.py <k> [fp64|fp32]"""
import os, sys, time; sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
import recommendersystems_b200 as rs, bench
from recommendersystems_b200.rwr import run_fixed
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
uid = [rs.Comm.unique_id() if rank == 0 else None]; dist.broadcast_object_list(uid, src=0)
comm = rs.Comm(rank, world, uid[0], device=local)
k = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
prec = rs.FP32 if len(sys.argv) > 2 and sys.argv[2] == "fp32" else rs.FP64
spec = dict(bench.C4_SPEC)
for f in ("n_users", "n_items", "n_like", "n_friend"):
    spec[f] = int(spec[f] * k)
t0 = time.perf_counter()
g = rs.Graph.synthetic(spec, comm=comm); g.buildGraph()
torch.cuda.synchronize()
i = g.info()
setup = time.perf_counter() - t0
free, total = torch.cuda.mem_get_info()
seed = 12345
c = rs.widen_float(0.15)
m = run_fixed(g, [seed], c, 20, prec)
m.rerun([seed], c, 20)
dist.barrier()
ts = []
for _ in range(3):
    m.rerun([seed], c, 20); ts.append(m.info().iterate_ms / 20)
ids, sc, cnt = m.topk(10)
s = m.scores(0)
mass = float(s.sum())
tops = [None] * world
dist.all_gather_object(tops, (ids[0].tolist(), float(s[seed])))
vb = 8 if prec == rs.FP64 else 4
alg = i.nnz / world * (4 + vb) + 4 * (i.n_nodes / world + 1) + i.n_nodes * vb + i.n_nodes / world * vb
ms = float(np.mean(ts))
sys.stdout.write(f"rank {rank}: n {i.n_nodes} nnz {i.nnz} rows [{i.row_begin},{i.row_end}) x_blocks {i.x_blocks} {ms:.4f} ms/iteration "
                 f"-> {i.nnz / ms / 1e6:.1f} GTEPS, hbm frac {alg / (ms * 1e-3) / 1e9 / bench.measured_peak_gbs()[0]:.3f}; synth {i.synth_ms:.0f} ms "
                 f"build {i.build_ms:.0f} ms setup {setup:.1f} s; handle {i.device_bytes / 1e9:.1f} GB, device in use {(total - free) / 1e9:.1f} GB; "
                 f"mass/N {mass / i.n_nodes:.12f}; same top-10 on every rank {all(t == tops[0] for t in tops)}\n")
sys.stdout.flush()
m.close(); g.close(); comm.close()
