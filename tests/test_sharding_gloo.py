"""N > 1 host logic on CPU: seed sharding + result gathering over torch.distributed (gloo, world_size 2).  The compute
inside each rank is the CPU oracle standing in for the GPU path (this file tests the plumbing, not the kernels)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank: int, world: int, port: int, out_dir: str):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import oracle as O
    from recommendersystems_b200.sharding import recommend_sharded, shard_seeds

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = dict(seed=77, n_users=300, n_items=2500, n_third=40, authorship_per_mille=800, n_like=9000, n_friend=2000,
                n_follow=100, n_mention=50, undefined_per_mille=20, scramble=1, p1_byte=61, reserved=0)
    links = O.synth_generate(spec)
    og = O.OracleGraph(links["node_id"], links["node_type"], links["src"], links["dst"], links["etype"], links["w"])
    assert og.build() == 0
    deg = np.bincount(links["src"], minlength=og.n)
    seeds = np.flatnonzero(deg[:300] > 0)[:11].astype(np.int32)          # 11 seeds: uneven blocks (5 + 6)
    k = 5

    def batch(local):
        ids = np.zeros((len(local), k), np.int64); sc = np.zeros((len(local), k)); cnt = np.zeros(len(local), np.int32)
        for i, s in enumerate(local):
            a, b = og.recommend(int(s), 0.15, 6, top_n=k)
            cnt[i] = len(a); ids[i, :len(a)] = a; sc[i, :len(a)] = b
        return ids, sc, cnt

    ids, sc, cnt = recommend_sharded(batch, seeds, k, dist)
    mine = shard_seeds(seeds, rank, world)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ids=ids, sc=sc, cnt=cnt, mine=mine)
    dist.barrier()
    dist.destroy_process_group()


def test_seed_sharding_world_size_2(tmp_path):
    import torch.multiprocessing as mp
    from recommendersystems_b200.sharding import shard_bounds
    assert shard_bounds(11, 2) == [0, 5, 11] and shard_bounds(1024, 8)[3] == 384 and shard_bounds(3, 4) == [0, 0, 1, 2, 3]
    port = 29700 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    # every rank holds the lists of all seeds, identical, in seed order; the blocks partition the seed list
    for key in ("ids", "sc", "cnt"):
        assert np.array_equal(r0[key], r1[key])
    assert len(r0["cnt"]) == 11 and len(r0["mine"]) + len(r1["mine"]) == 11
    # against the unsharded run
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from recommendersystems_b200.sharding import recommend_sharded
    spec = dict(seed=77, n_users=300, n_items=2500, n_third=40, authorship_per_mille=800, n_like=9000, n_friend=2000,
                n_follow=100, n_mention=50, undefined_per_mille=20, scramble=1, p1_byte=61, reserved=0)
    links = O.synth_generate(spec)
    og = O.OracleGraph(links["node_id"], links["node_type"], links["src"], links["dst"], links["etype"], links["w"])
    assert og.build() == 0
    deg = np.bincount(links["src"], minlength=og.n)
    seeds = np.flatnonzero(deg[:300] > 0)[:11]
    for i, s in enumerate(seeds):
        a, b = og.recommend(int(s), 0.15, 6, top_n=5)
        assert r0["ids"][i, :r0["cnt"][i]].tolist() == a.tolist()
