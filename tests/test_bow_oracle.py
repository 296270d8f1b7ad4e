"""Bag-of-words transform: the oracle restatement against the reference's own DBoW2 sources
compiled unmodified (oracle/_ref/libref_bow.so), on synthetic vocabularies in the reference's text format."""
import os

import numpy as np
import pytest

from dani_slam_b200 import synth
from oracle import oracle, ref_binding

needs_ref = pytest.mark.skipif(not ref_binding.bow_available(), reason="oracle/_ref/libref_bow.so not built (needs /root/reference)")

CASES = [
    dict(k=10, L=3, seed=1),                                   # ORBvoc shape, shallower
    dict(k=10, L=4, seed=2, stop_frac=0.2),                    # many stopped words
    dict(k=4, L=6, seed=3, flips=6),                           # deep, close siblings → distance ties
    dict(k=7, L=5, seed=4, ragged=0.3, min_leaf_level=2),      # early leaves (never above the levelsup level used)
    dict(k=10, L=3, seed=5, scoring=1),                        # L2 norm
    dict(k=10, L=3, seed=6, scoring=5),                        # DOT_PRODUCT: divide by size instead of normalising
    dict(k=10, L=3, seed=7, weighting=2),                      # IDF: addIfNotExist
    dict(k=10, L=3, seed=8, weighting=3, scoring=1),           # BINARY + L2
    dict(k=10, L=3, seed=9, weighting=1),                      # TF
]


def _same(a, b, keys):
    for k in keys:
        assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, k
        assert np.array_equal(a[k], b[k]), k      # doubles included: bit-exact


@needs_ref
@pytest.mark.parametrize("case", CASES, ids=lambda c: "k%d_L%d_s%d" % (c["k"], c["L"], c["seed"]))
def test_oracle_matches_reference_dbow2(case, tmp_path):
    voc = synth.vocabulary(**case)
    path = os.path.join(tmp_path, "voc.txt")
    synth.write_vocabulary_text(voc, path)
    ref = ref_binding.Vocabulary(path)
    orc_txt = oracle.Vocabulary(path=path)
    orc_arr = oracle.Vocabulary(voc=voc)
    assert ref.n_words == orc_txt.n_words == orc_arr.n_words == int(voc["is_leaf"].sum())
    levelsup = case["L"] - case.get("min_leaf_level", case["L"]) if case.get("ragged") else min(4, case["L"])
    for n, seed in ((1000, 0), (1, 1), (0, 2), (2500, 3)):
        q = synth.vocabulary_queries(voc, n, seed=seed)
        r = ref.transform(q, levelsup)
        for o in (orc_txt.transform(q, levelsup), orc_arr.transform(q, levelsup)):
            _same(r, o, ["word_id", "bow_ids", "bow_vals", "fv_nodes", "fv_off", "fv_idx"])
    # every levelsup on the full-depth vocabularies
    if not case.get("ragged"):
        q = synth.vocabulary_queries(voc, 500, seed=9)
        for lu in range(0, case["L"] + 2):
            _same(ref.transform(q, lu), orc_arr.transform(q, lu), ["bow_ids", "bow_vals", "fv_nodes", "fv_off", "fv_idx"])


@needs_ref
def test_l1_score_matches_reference(tmp_path):
    voc = synth.vocabulary(k=10, L=3, seed=21)
    path = os.path.join(tmp_path, "voc.txt")
    synth.write_vocabulary_text(voc, path)
    ref = ref_binding.Vocabulary(path)
    orc = oracle.Vocabulary(voc=voc)
    a = orc.transform(synth.vocabulary_queries(voc, 800, seed=1))
    for seed in (1, 2, 3):
        b = orc.transform(synth.vocabulary_queries(voc, 700, seed=seed))
        s_ref = ref.score(a["bow_ids"], a["bow_vals"], b["bow_ids"], b["bow_vals"])
        s_orc = oracle.bow_score_l1(a["bow_ids"], a["bow_vals"], b["bow_ids"], b["bow_vals"])
        assert s_ref == s_orc
    assert oracle.bow_score_l1(a["bow_ids"], a["bow_vals"], a["bow_ids"], a["bow_vals"]) == pytest.approx(1.0, abs=1e-12)


def test_oracle_properties():
    voc = synth.vocabulary(k=10, L=3, seed=31)
    orc = oracle.Vocabulary(voc=voc)
    q = synth.vocabulary_queries(voc, 1200, seed=5)
    t = orc.transform(q, 2)
    kept = np.asarray(voc["weight"])[np.flatnonzero(voc["is_leaf"])][t["word_id"]] > 0
    assert np.all(np.diff(t["bow_ids"].astype(np.int64)) > 0) and np.all(np.diff(t["fv_nodes"].astype(np.int64)) > 0)
    assert abs(t["bow_vals"].sum() - 1.0) < 1e-12                        # L1 normalised
    assert set(t["bow_ids"]) == set(t["word_id"][kept])
    assert sorted(t["fv_idx"]) == list(np.flatnonzero(kept))             # every kept feature exactly once
    for a, b in zip(t["fv_off"][:-1], t["fv_off"][1:]):
        assert np.all(np.diff(t["fv_idx"][a:b].astype(np.int64)) > 0)    # feature order inside a node
    # node ids are the ancestors at level L - levelsup = 1: children of the root
    assert set(t["fv_nodes"]) <= set(np.flatnonzero(voc["parent"] == 0) + 1)


# ---- golden vectors produced by the reference's own DBoW2 (tools/gen_golden_bow.py); usable without oracle/_ref ----
import glob
import hashlib

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bow_*.npz")))


def golden_case(path):
    g = np.load(path)
    vargs = eval(str(g["vargs"]))          # a dict literal written by the generator script
    voc = synth.vocabulary(**vargs)
    q = synth.vocabulary_queries(voc, int(g["nq"]), seed=vargs["seed"] + 1)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert sha(np.concatenate([voc["parent"].view(np.uint8), voc["is_leaf"], voc["desc"].ravel(), voc["weight"].view(np.uint8)])) == str(g["voc_sha"]), "vocabulary generator drifted"
    assert sha(q) == str(g["q_sha"]), "query generator drifted"
    return g, voc, q, int(g["levelsup"])


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[4:-4])
def test_oracle_matches_golden_dbow2_vectors(path):
    g, voc, q, levelsup = golden_case(path)
    o = oracle.Vocabulary(voc=voc).transform(q, levelsup)
    for k in ("word_id", "bow_ids", "bow_vals", "fv_nodes", "fv_off", "fv_idx"):
        assert np.array_equal(o[k], g[k]), k


def test_golden_set_is_present():
    assert len(GOLDEN) >= 5
