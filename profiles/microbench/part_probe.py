"""Row-partitioned iteration time with and without the exchange (torchrun, one process per GPU)."""
import os, sys; sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
import recommendersystems_b200 as rs, bench
from recommendersystems_b200.rwr import run_fixed
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
uid = [rs.Comm.unique_id() if rank == 0 else None]; dist.broadcast_object_list(uid, src=0)
comm = rs.Comm(rank, world, uid[0], device=local)
# argument: a scale factor of the C2 graph, or "c4" = BASELINE configs[3] (50 M nodes, ~2 B links)
if len(sys.argv) > 1 and sys.argv[1] == "c4":
    spec = dict(seed=20260104, n_users=5_000_000, n_items=45_000_000, n_third=0, authorship_per_mille=1000, n_like=700_000_000,
                n_friend=200_000_000, n_follow=0, n_mention=0, undefined_per_mille=0, scramble=1, p1_byte=61)
else:
    spec = bench.scaled_spec(float(sys.argv[1]) if len(sys.argv) > 1 else 1.0)
import time
t0 = time.perf_counter()
g = rs.Graph.synthetic(spec, comm=comm, hub_entries=int(os.environ.get("HUB", "-1"))); g.buildGraph()
torch.cuda.synchronize()
if rank == 0: sys.stdout.write(f"setup wall {time.perf_counter() - t0:.1f} s synth {g.info().synth_ms:.0f} ms build {g.info().build_ms:.0f} ms\n"); sys.stdout.flush()
i = g.info()
c = rs.widen_float(0.15)
seed = 12345
precs = ((rs.FP64, 'fp64'),) if (len(sys.argv) > 1 and sys.argv[1] == 'c4') else ((rs.FP64, 'fp64'), (rs.FP32, 'fp32'))
for prec, pn in precs:
    m = run_fixed(g, [seed], c, 20, prec)
    for _ in range(2): m.rerun([seed], c, 20)
    dist.barrier()
    ts = []
    for _ in range(5):
        m.rerun([seed], c, 20); ts.append(m.info().iterate_ms / 20)
    sys.stdout.write(f"rank {rank} rows [{i.row_begin},{i.row_end}) nnz {i.nnz} {pn}: {np.mean(ts):.4f} ms/iteration skip={os.environ.get('RWR_DIST_SKIP')} nop2p={os.environ.get('RWR_DIST_NO_P2P')} build {i.build_ms:.0f} ms dev {i.device_bytes/1e9:.1f} GB\n"); sys.stdout.flush()
    m.close()
g.close(); comm.close()
