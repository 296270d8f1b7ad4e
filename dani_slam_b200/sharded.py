"""Multi-GPU plumbing for the two workloads of SURVEY.md §8(e): one process per GPU, `torch.distributed`.

* frame-batch extraction: frames are independent → contiguous split of the batch, NO collective.
* DB-sharded Hamming kNN: the descriptor database is split into contiguous index ranges (so that
  "lower train index wins ties" stays a lexicographic minimum on (dist, global idx)); every rank computes
  its local top-2 with global indices, ONE all-gather of the packed record {idx[nq×2], dist[nq×2]} int32 follows (NCCL over NVLink
  on GPUs, gloo in the CPU tests), then every rank merges the G×2 candidates per query.

torch is used only for device memory, streams and the collective; the compute is liborbx's kernels
(`local_top2` / `merge` are injected so the CPU tests can exercise the plumbing with gloo).
"""
from __future__ import annotations

from typing import Callable, Tuple


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of `n` items owned by `rank` out of `world` (sizes differ by at most one)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_top2(rec, group=None):
    """ONE all-gather of the per-rank local top-2 record `rec` = [2, nq, 2] int32 (plane 0: global indices, plane 1:
    distances) → [world, 2, nq, 2]."""
    import torch
    import torch.distributed as td
    world = td.get_world_size(group)
    rec = rec.contiguous()
    # output = concatenation along dim 0 (the layout both NCCL and gloo accept), viewed as [world, 2, nq, 2]
    out = torch.empty((world * rec.shape[0],) + tuple(rec.shape[1:]), dtype=rec.dtype, device=rec.device)
    td.all_gather_into_tensor(out, rec, group=group)
    return out.view(world, *rec.shape)


def sharded_knn2(query, db_shard, shard_lo: int, local_top2: Callable, merge: Callable, group=None):
    """Global top-2 of `query` against a database split across the ranks of `group`.

    local_top2(query, db_shard, idx_base) -> rec[2, nq, 2] int32: plane 0 = GLOBAL indices (idx_base added), plane 1 =
    distances, missing neighbours = (-1, INT32_MAX);  merge(rec_all[G, 2, nq, 2]) -> (idx[nq,2], dist[nq,2]).
    The result is bit-identical to the unsharded search, ties included.
    """
    rec = local_top2(query, db_shard, shard_lo)
    return merge(allgather_top2(rec, group))


class CudaShardedMatcher:
    """DB-sharded kNN on this rank's GPU: liborbx kernels + one NCCL all-gather (torch.distributed)."""

    def __init__(self, device_index: int):
        import torch
        from . import orbx
        self.torch = torch
        self.dev = torch.device("cuda", device_index)
        self.m = orbx.ORBmatcher(0.7, True, device=device_index)
        self._ext = torch.cuda.ExternalStream(self.m.stream(), device=self.dev)

    def local_top2(self, d_query, d_db, idx_base: int):
        torch = self.torch
        nq, ndb = d_query.shape[0], d_db.shape[0]
        rec = torch.empty((2, nq, 2), dtype=torch.int32, device=self.dev)
        self._ext.wait_stream(torch.cuda.current_stream(self.dev))
        self.m.knn2_device(d_query.data_ptr(), nq, d_db.data_ptr(), ndb, idx_base, rec[0].data_ptr(), rec[1].data_ptr())
        torch.cuda.current_stream(self.dev).wait_stream(self._ext)
        return rec

    def merge(self, rec_all):
        torch = self.torch
        G, nq = rec_all.shape[0], rec_all.shape[2]
        idx = torch.empty((nq, 2), dtype=torch.int32, device=self.dev)
        dist = torch.empty((nq, 2), dtype=torch.int32, device=self.dev)
        self._ext.wait_stream(torch.cuda.current_stream(self.dev))
        self.m.merge_packed_device(rec_all.data_ptr(), G, nq, idx.data_ptr(), dist.data_ptr())
        torch.cuda.current_stream(self.dev).wait_stream(self._ext)
        return idx, dist

    def knn2(self, d_query, d_db_shard, shard_lo: int, group=None):
        return sharded_knn2(d_query, d_db_shard, shard_lo, self.local_top2, self.merge, group)

    # ---- the C-ABI path: orbx_knn2_sharded does the scan, the NCCL all-gather and the merge inside liborbx ----
    def init_comm(self, group=None):
        """Builds this rank's orbx communicator; the 128-byte NCCL id travels over torch.distributed (plumbing only)."""
        import torch.distributed as td
        from . import orbx
        torch = self.torch
        rank, world = td.get_rank(group), td.get_world_size(group)
        uid = torch.zeros(128, dtype=torch.uint8, device=self.dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(orbx.Comm.unique_id()), dtype=torch.uint8))
        td.broadcast(uid, src=0, group=group)
        self.comm = orbx.Comm(world, rank, bytes(uid.cpu().numpy().tobytes()), self.dev.index)
        return self.comm

    def knn2_cabi(self, d_query, d_db_shard, shard_lo: int):
        torch = self.torch
        nq = d_query.shape[0]
        idx = torch.empty((nq, 2), dtype=torch.int32, device=self.dev)
        dist = torch.empty((nq, 2), dtype=torch.int32, device=self.dev)
        self._ext.wait_stream(torch.cuda.current_stream(self.dev))
        self.m.knn2_sharded(self.comm, d_query.data_ptr(), nq, d_db_shard.data_ptr(), d_db_shard.shape[0], shard_lo, idx.data_ptr(), dist.data_ptr())
        torch.cuda.current_stream(self.dev).wait_stream(self._ext)
        return idx, dist
