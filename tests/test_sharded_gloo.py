"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: DB-sharded kNN with the top-2 all-gather,
and per-rank frame-batch sharding.  The local top-2 / merge kernels are replaced by the CPU oracle here —
this file tests the plumbing (shard ranges, global indices, collective, merge order), not the kernels."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INT_MAX = 2**31 - 1


def _merge_np(idx_all, dist_all):
    G, nq, _ = idx_all.shape
    idx = np.full((nq, 2), -1, np.int32)
    dist = np.full((nq, 2), INT_MAX, np.int32)
    for i in range(nq):
        c = [(int(dist_all[g, i, k]), int(idx_all[g, i, k])) for g in range(G) for k in range(2) if idx_all[g, i, k] >= 0]
        c.sort()
        for k, (d, j) in enumerate(c[:2]):
            idx[i, k], dist[i, k] = j, d
    return idx, dist


def _worker(rank, world, port, ndb, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dani_slam_b200 import sharded, synth
        from oracle import oracle
        q, db = synth.knn_case(40, ndb, seed=11, planted_frac=0.2)
        lo, hi = sharded.shard_bounds(ndb, rank, world)

        def local_top2(qq, shard, base):
            i, d = oracle.knn2(qq, shard)
            i = np.where(i >= 0, i + base, -1).astype(np.int32)
            return torch.from_numpy(np.stack([i, d.astype(np.int32)]))     # the packed record [2, nq, 2]

        def merge(rec_all):
            a = rec_all.numpy()
            i, d = _merge_np(a[:, 0], a[:, 1])
            return torch.from_numpy(i), torch.from_numpy(d)

        idx, dist = sharded.sharded_knn2(q, db[lo:hi], lo, local_top2, merge)
        ridx, rdist = oracle.knn2(q, db)
        ok = np.array_equal(idx.numpy(), ridx) and np.array_equal(dist.numpy(), rdist)
        # frame-batch sharding: no collective, every frame processed exactly once
        flo, fhi = sharded.shard_bounds(37, rank, world)
        mine = torch.zeros(37, dtype=torch.int32)
        mine[flo:fhi] = 1
        td.all_reduce(mine)
        ok = ok and bool((mine == 1).all())
        out[rank] = int(ok)
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("ndb", [1001, 3, 1])
def test_sharded_knn_two_ranks_equals_unsharded(ndb):
    from oracle import oracle
    oracle.build()
    world = 2
    port = 29500 + (os.getpid() % 2000) + ndb % 7
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    procs = [ctx.Process(target=_worker, args=(r, world, port, ndb, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1] * world
