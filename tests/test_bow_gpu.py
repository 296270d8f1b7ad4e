"""Bag-of-words transform on the GPU through the C ABI: bit-exact against the oracle, which
tests/test_bow_oracle.py pins to the reference's own DBoW2."""
import os

import numpy as np
import pytest

from dani_slam_b200 import orbx, synth
from oracle import oracle

pytestmark = pytest.mark.gpu

KEYS = ["word_id", "node_id", "bow_ids", "bow_vals", "fv_nodes", "fv_off", "fv_idx"]

CASES = [
    dict(k=10, L=3, seed=1),
    dict(k=10, L=4, seed=2, stop_frac=0.2),
    dict(k=4, L=6, seed=3, flips=6),
    dict(k=7, L=5, seed=4, ragged=0.3, min_leaf_level=2),
    dict(k=20, L=2, seed=12),                                   # widest fan-out the text format accepts
    dict(k=10, L=3, seed=5, scoring=1),
    dict(k=10, L=3, seed=6, scoring=5),
    dict(k=10, L=3, seed=7, weighting=2),
    dict(k=10, L=3, seed=8, weighting=3, scoring=1),
    dict(k=10, L=3, seed=9, weighting=1),
]


def _same(a, b):
    for k in KEYS:
        assert a[k].shape == b[k].shape, k
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("case", CASES, ids=lambda c: "k%d_L%d_s%d" % (c["k"], c["L"], c["seed"]))
def test_transform_matches_oracle(case):
    voc = synth.vocabulary(**case)
    orc = oracle.Vocabulary(voc=voc)
    dev = orbx.ORBVocabulary().from_nodes(voc)
    assert dev.size() == orc.n_words
    for n, seed in ((1000, 0), (1, 1), (0, 2), (2500, 3), (33, 4)):
        q = synth.vocabulary_queries(voc, n, seed=seed)
        for lu in (0, 1, 2, 4, case["L"], case["L"] + 1):
            _same(orc.transform(q, lu), dev.transform(q, lu))


def test_text_loader_and_errors(tmp_path):
    voc = synth.vocabulary(k=10, L=3, seed=41)
    path = os.path.join(tmp_path, "voc.txt")
    synth.write_vocabulary_text(voc, path)
    with open(path, "a") as f:
        f.write("\n\n")                       # trailing blank lines are ignored
    dev = orbx.ORBVocabulary()
    assert dev.loadFromTextFile(path)
    orc = oracle.Vocabulary(path=path)
    q = synth.vocabulary_queries(voc, 1500, seed=2)
    _same(orc.transform(q, 4), dev.transform(q, 4))
    info = dev.info()
    assert info["k"] == 10 and info["L"] == 3 and info["n_words"] == int(voc["is_leaf"].sum()) and info["n_nodes"] == len(voc["parent"]) + 1
    bad = os.path.join(tmp_path, "bad.txt")
    open(bad, "w").write("99 3 0 0\n0 1 " + " ".join(["0"] * 32) + " 1.0")
    assert not orbx.ORBVocabulary().loadFromTextFile(bad)
    assert not orbx.ORBVocabulary().loadFromTextFile(os.path.join(tmp_path, "missing.txt"))
    with pytest.raises(orbx.OrbxError):
        orbx.ORBVocabulary().from_nodes(dict(voc, parent=np.full_like(voc["parent"], 5)))   # parents must precede children


def test_batch_device_on_extracted_descriptors():
    """Extractor output → transform, without leaving the device: per-image results equal the oracle's on the same rows."""
    import torch
    B, H, W, cap = 6, 240, 320, 640
    frames = np.stack([synth.parity_frame(s, W, H) if s % 2 else synth.throughput_frame(s, W, H) for s in range(B)])
    ex = orbx.ORBextractor(500, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
    d_img = torch.from_numpy(frames).cuda()
    d_kps = torch.zeros(B * cap * 7, dtype=torch.float32, device="cuda")
    d_desc = torch.zeros(B * cap * 32, dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(B, dtype=torch.int32, device="cuda"); d_mono = torch.zeros(B, dtype=torch.int32, device="cuda")
    ex.extract_batch_device(d_img.data_ptr(), H * W, B, H, W, W, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_n.data_ptr(), d_mono.data_ptr(), (0, 0))
    ex.sync()
    n = d_n.cpu().numpy()
    assert n.min() > 50
    voc = synth.vocabulary(k=10, L=4, seed=77)
    dev = orbx.ORBVocabulary().from_nodes(voc)
    orc = oracle.Vocabulary(voc=voc)
    i32, u32 = torch.int32, torch.int32       # same width; viewed as uint32 on the host
    word = torch.zeros(B * cap, dtype=u32, device="cuda"); node = torch.zeros_like(word)
    bow_ids = torch.zeros_like(word); bow_vals = torch.zeros(B * cap, dtype=torch.float64, device="cuda")
    fv_nodes = torch.zeros_like(word); fv_idx = torch.zeros_like(word); fv_off = torch.zeros(B * (cap + 1), dtype=i32, device="cuda")
    n_bow = torch.zeros(B, dtype=i32, device="cuda"); n_fv = torch.zeros(B, dtype=i32, device="cuda")
    dev.transform_batch_device(d_desc.data_ptr(), cap * 32, d_n.data_ptr(), B, cap, 2, word.data_ptr(), node.data_ptr(), bow_ids.data_ptr(),
                               bow_vals.data_ptr(), n_bow.data_ptr(), fv_nodes.data_ptr(), fv_off.data_ptr(), fv_idx.data_ptr(), n_fv.data_ptr())
    dev.sync()
    desc = d_desc.cpu().numpy().reshape(B, cap, 32)
    word, node, bow_ids, fv_nodes, fv_idx = [t.cpu().numpy().view(np.uint32).reshape(B, cap) for t in (word, node, bow_ids, fv_nodes, fv_idx)]
    bow_vals = bow_vals.cpu().numpy().reshape(B, cap); fv_off = fv_off.cpu().numpy().reshape(B, cap + 1)
    n_bow, n_fv = n_bow.cpu().numpy(), n_fv.cpu().numpy()
    for b in range(B):
        o = orc.transform(desc[b, :n[b]], 2)
        got = dict(word_id=word[b, :n[b]], node_id=node[b, :n[b]], bow_ids=bow_ids[b, :n_bow[b]], bow_vals=bow_vals[b, :n_bow[b]],
                   fv_nodes=fv_nodes[b, :n_fv[b]], fv_off=fv_off[b, :n_fv[b] + 1], fv_idx=fv_idx[b, :fv_off[b, n_fv[b]]])
        _same(o, got)


def test_orbvoc_sized_tree_properties():
    """k = 10, L = 5 (111 k nodes; the full ORBvoc has L = 6): size-independent checks plus a sampled oracle comparison."""
    voc = synth.vocabulary(k=10, L=5, seed=5, flips=48)
    dev = orbx.ORBVocabulary().from_nodes(voc)
    q = synth.vocabulary_queries(voc, 8000, seed=1)
    t = dev.transform(q, 4)
    assert np.all(np.diff(t["bow_ids"].astype(np.int64)) > 0) and np.all(np.diff(t["fv_nodes"].astype(np.int64)) > 0)
    assert abs(t["bow_vals"].sum() - 1.0) < 1e-9
    w_leaf = np.asarray(voc["weight"])[np.flatnonzero(voc["is_leaf"])]
    kept = w_leaf[t["word_id"]] > 0
    assert sorted(t["fv_idx"]) == list(np.flatnonzero(kept))
    # word ids do not depend on what else is in the batch (idempotence under permutation)
    perm = np.random.default_rng(0).permutation(len(q))
    t2 = dev.transform(q[perm], 4)
    assert np.array_equal(t2["word_id"], t["word_id"][perm])
    assert np.array_equal(t2["bow_ids"], t["bow_ids"])
    orc = oracle.Vocabulary(voc=voc)
    _same(orc.transform(q[:2000], 4), dev.transform(q[:2000], 4))


def test_vocabulary_shared_between_threads_and_capacity_error():
    """The reference shares one const vocabulary between Tracking, LocalMapping and LoopClosing threads: concurrent
    transform calls on one handle must give the single-threaded results.  More than 16384 descriptors per image is a
    capacity error, not a crash."""
    import threading
    voc = synth.vocabulary(k=10, L=4, seed=51)
    dev = orbx.ORBVocabulary().from_nodes(voc)
    orc = oracle.Vocabulary(voc=voc)
    qs = [synth.vocabulary_queries(voc, 700 + 37 * i, seed=i) for i in range(8)]
    want = [orc.transform(q, 4) for q in qs]
    got = [None] * len(qs)

    def work(i):
        for _ in range(5):
            got[i] = dev.transform(qs[i], 4)

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(qs))]
    [t.start() for t in th]
    [t.join() for t in th]
    for w, g in zip(want, got):
        _same(w, g)
    with pytest.raises(orbx.OrbxError) as e:
        dev.transform(synth.vocabulary_queries(voc, 20000, seed=3), 4)
    assert e.value.code == orbx.ERR_CAPACITY
    _same(want[0], dev.transform(qs[0], 4))          # the handle stays usable


def test_transform_matches_golden_dbow2_vectors():
    """Golden vectors written by the reference's own DBoW2 (tools/gen_golden_bow.py)."""
    from test_bow_oracle import GOLDEN, golden_case
    assert len(GOLDEN) >= 5
    for path in GOLDEN:
        g, voc, q, levelsup = golden_case(path)
        t = orbx.ORBVocabulary().from_nodes(voc).transform(q, levelsup)
        for k in ("word_id", "bow_ids", "bow_vals", "fv_nodes", "fv_off", "fv_idx"):
            assert np.array_equal(t[k], g[k]), (os.path.basename(path), k)
