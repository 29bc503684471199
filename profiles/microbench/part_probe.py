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
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
g = rs.Graph.synthetic(bench.scaled_spec(scale), comm=comm); g.buildGraph()
i = g.info()
c = rs.widen_float(0.15)
seed = 12345
for prec, pn in ((rs.FP64, "fp64"), (rs.FP32, "fp32")):
    m = run_fixed(g, [seed], c, 20, prec)
    for _ in range(2): m.rerun([seed], c, 20)
    dist.barrier()
    ts = []
    for _ in range(5):
        m.rerun([seed], c, 20); ts.append(m.info().iterate_ms / 20)
    print(f"rank {rank} rows [{i.row_begin},{i.row_end}) nnz {i.nnz} {pn}: {np.mean(ts):.4f} ms/iteration skip={os.environ.get('RWR_DIST_SKIP')} build {i.build_ms:.0f} ms dev {i.device_bytes/1e9:.1f} GB", flush=True)
    m.close()
g.close(); comm.close()
