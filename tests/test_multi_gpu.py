"""Multi-GPU entry points of the C ABI on real hardware (needs >= 2 CUDA devices; skipped on a single-GPU box):
orbx_knn2_sharded_all (DB-sharded kNN, one NCCL all-gather, merge) == the unsharded oracle scan, and orbx_extract_batch_multi
(frame batch split over devices, no collective) == the single-device result."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ndev(orbx_mod):
    n = orbx_mod.lib().orbx_device_count()
    if n < 2:
        pytest.skip("needs at least 2 CUDA devices")
    return min(n, 8)


@pytest.mark.parametrize("nq,ndb", [(2000, 1_000_003), (77, 5), (300, 40_000)])
def test_knn2_sharded_all_equals_unsharded(orbx_mod, oracle_mod, ndev, nq, ndb):
    import torch
    from dani_slam_b200 import sharded, synth
    q, db = synth.knn_case(nq, ndb, seed=nq + ndb, planted_frac=0.05, dup_rows=4)
    ridx, rdist = oracle_mod.knn2(q, db, nthreads=8)
    ms = [orbx_mod.ORBmatcher(0.7, True, device=d) for d in range(ndev)]
    cs = orbx_mod.Comm.create_all(list(range(ndev)))
    assert [c.rank for c in cs] == list(range(ndev)) and cs[0].world == ndev
    dq, ddb, di, dd, los, ns = [], [], [], [], [], []
    for d in range(ndev):
        lo, hi = sharded.shard_bounds(ndb, d, ndev)
        dev = torch.device("cuda", d)
        dq.append(torch.from_numpy(q).to(dev)); ddb.append(torch.from_numpy(db[lo:hi].copy()).to(dev))
        di.append(torch.full((nq, 2), -7, dtype=torch.int32, device=dev)); dd.append(torch.full((nq, 2), -7, dtype=torch.int32, device=dev))
        los.append(lo); ns.append(hi - lo)
    for d in range(ndev):
        torch.cuda.synchronize(d)
    for _ in range(2):                                                     # twice: buffers and communicators are reusable
        orbx_mod.knn2_sharded_all(ms, cs, [t.data_ptr() for t in dq], nq, [t.data_ptr() if t.numel() else 0 for t in ddb], ns, los,
                                  [t.data_ptr() for t in di], [t.data_ptr() for t in dd])
        for m in ms:
            m.sync()
        for d in range(ndev):                                              # every rank holds the global result
            assert np.array_equal(di[d].cpu().numpy(), ridx) and np.array_equal(dd[d].cpu().numpy(), rdist), d
    for c in cs:
        c.close()


def test_extract_batch_multi_equals_single_device(orbx_mod, ndev):
    from dani_slam_b200 import synth
    W, H, nf, B = 640, 480, 1000, 37
    frames = np.stack([synth.throughput_frame(100 + i, W, H) for i in range(B)])
    cap = nf + 24
    exs = [orbx_mod.ORBextractor(nf, 1.2, 8, 20, 7, device=d, max_width=W, max_height=H, max_batch=B) for d in range(ndev)]
    k, de, n, mono = orbx_mod.extract_batch_multi(exs, frames, cap)
    k1, de1, n1, mono1 = orbx_mod.extract_batch_multi(exs[:1], frames, cap)
    assert np.array_equal(n, n1) and np.array_equal(mono, mono1) and n.min() >= nf
    for b in range(B):
        assert k[b, : n[b]].tobytes() == k1[b, : n1[b]].tobytes() and np.array_equal(de[b, : n[b]], de1[b, : n1[b]]), b
