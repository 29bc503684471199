"""Worker of the row-partitioned parity check: one process per GPU (torchrun or the spawn in test_gpu_partitioned.py).

Every rank builds the same graph row-partitioned over all ranks, runs fixed and threshold iterations and top-k, and
compares with the CPU oracle on the same links.  Prints one `PARTITIONED OK` line per rank on success."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    import oracle as O
    import recommendersystems_b200 as rs
    from recommendersystems_b200.rwr import run_fixed, run_threshold

    mode = sys.argv[1] if len(sys.argv) > 1 else "overlapped"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    uid = [rs.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = rs.Comm(rank, world, uid[0], device=local)

    c = rs.widen_float(0.15)
    for spec, valued in ((dict(seed=11, n_users=3_000, n_items=40_000, n_third=500, authorship_per_mille=700, n_like=150_000,
                               n_friend=30_000, n_follow=2_000, n_mention=0, undefined_per_mille=20, scramble=1, p1_byte=50), False),
                         (dict(seed=12, n_users=2_000, n_items=20_000, n_third=300, authorship_per_mille=700, n_like=80_000,
                               n_friend=20_000, n_follow=1_000, n_mention=1_500, undefined_per_mille=20, scramble=1, p1_byte=50), True),
                         # ~3.4 M links: every stream block spans many tiles, so the SpMV of the overlapped exchange really
                         # runs ahead of the slices that are still arriving
                         (dict(seed=13, n_users=40_000, n_items=360_000, n_third=0, authorship_per_mille=800, n_like=1_400_000,
                               n_friend=300_000, n_follow=0, n_mention=0, undefined_per_mille=0, scramble=1, p1_byte=61), False)):
        links = O.synth_generate(spec)          # the generator is deterministic: the oracle's CPU copy of the same graph
        if valued:
            # the host-array entry point: every rank is handed the whole list and uploads the links of the sources it owns
            g = rs.Graph.from_arrays(links["node_id"], links["node_type"], links["src"], links["dst"], links["etype"], links["w"], comm=comm)
        else:
            g = rs.Graph.synthetic(spec, comm=comm)
        g.buildGraph()
        info = g.info()
        assert info.n_ranks == world and 0 <= info.row_begin <= info.row_end <= info.n_nodes
        want_blocks = {"overlapped": world, "overlapped_slow": world, "peer_stores": 1, "nccl": 1, "peer_stores_blocked": 3, "replicated_build": 1}[mode]
        assert info.x_blocks == want_blocks, (mode, info.x_blocks)
        assert info.layout == (rs._native.LAYOUT_VALUED if valued else rs._native.LAYOUT_INDEX)
        rows = torch.tensor([info.row_end - info.row_begin])
        dist.all_reduce(rows)
        assert int(rows) == info.n_nodes, "the slices must cover every row exactly once"
        held = torch.tensor([info.n_links_raw])
        dist.all_reduce(held)
        assert int(held) == len(links["src"]) * (world if mode == "replicated_build" else 1)
        assert np.array_equal(g.degrees(raw=True), np.bincount(links["src"], minlength=info.n_nodes))
        og = O.OracleGraph(links["node_id"], links["node_type"], links["src"], links["dst"], links["etype"], links["w"])
        assert og.build() == 0
        assert og.nnz() == info.nnz
        local = g.export_links()                # a partitioned build keeps the links of the owned sources only
        if mode != "replicated_build":
            assert 0 < len(local["src"]) < len(links["src"])
        deg = np.bincount(links["src"], minlength=og.n)
        seeds = [int(np.argmax(deg)), int(np.flatnonzero(deg[:spec["n_users"]] > 3)[5])]
        for seed in seeds:
            for prec, tol in ((rs.FP64, 1e-12), (rs.FP32, None)):
                r = run_fixed(g, [seed], c, 9, prec)
                got = r.scores(0)
                want, _ = og.run(seed, c, n_iter=9)
                if tol is not None:
                    nz = want != 0
                    assert np.all(got[~nz] == 0) and (np.abs(got[nz] - want[nz]) / want[nz]).max() <= tol
                else:
                    assert np.abs(got / got.sum() - want / want.sum()).sum() <= 1e-6
                ids, sc, cnt = r.topk(10)
                oids, _ = og.recommend(seed, 0.15, 9, top_n=10)
                if prec == rs.FP64:
                    assert ids[0, :cnt[0]].tolist() == oids.tolist()
                r.close()
            res, it = run_threshold(g, [seed], c, 1e-9 * og.n, max_iter=300)
            _, want_it = og.run(seed, c, threshold=1e-9 * og.n)
            assert int(it[0]) == want_it, (int(it[0]), want_it)
            res.close()
            top = rs.Recommender(g).Recommendation(seed, 0.15, 9, 10)
            oids, _ = og.recommend(seed, 0.15, 9, top_n=10)
            assert [p[0] for p in top] == oids.tolist()
        g.close()
    dist.barrier()
    comm.close()
    dist.destroy_process_group()
    print(f"PARTITIONED OK rank {rank}/{world}", flush=True)


if __name__ == "__main__":
    main()
