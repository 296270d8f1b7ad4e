"""The C++ oracle end to end: against the committed golden vectors (made from real cv2 primitives by
tools/gen_golden.py) and, where cv2 is importable, live against oracle/cv2_oracle.py."""
import hashlib

import numpy as np
import pytest

from conftest import golden_cases, golden_frame, load_golden


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_golden(oracle_mod, name):
    g = load_golden(name)
    img = golden_frame(g)
    ex = oracle_mod.Extractor(int(g["nfeatures"]), 1.2, 8, 20, 7)
    rc, k, d, mono = ex.extract(img, rects=g["rects"], lap=tuple(int(v) for v in g["lap"]), cap=int(g["nfeatures"]) + 200)
    assert rc == 0
    for l in range(8):
        assert sha(ex.level(l, True)) == str(g["pyr_sha"][l]), f"pyramid level {l}"
        b = ex.blurred(l)
        assert (sha(b) if b is not None else "") == str(g["blur_sha"][l]), f"blurred level {l}"
        c = ex.candidates(l)
        assert sha(np.stack([c["x"], c["y"], c["response"]], axis=1).astype(np.float32)) == str(g["cand_sha"][l]), f"candidates {l}"
        s = ex.selected(l)
        assert sha(np.stack([s["x"], s["y"], s["response"]], axis=1).astype(np.float32)) == str(g["sel_sha"][l]), f"selected {l}"
    assert mono == int(g["mono"])
    assert k.tobytes() == g["kps"].tobytes()
    assert np.array_equal(d, g["desc"])


def test_oracle_matches_cv2_backed_restatement_live(oracle_mod):
    pytest.importorskip("cv2")
    from dani_slam_b200 import synth
    from oracle import cv2_oracle
    img = synth.parity_frame(21, 400, 300)
    rects = [(50, 40, 60, 60)]
    rc, k, d, mono = oracle_mod.Extractor(600, 1.2, 8, 20, 7).extract(img, rects=rects, lap=(100, 250))
    k2, d2, mono2 = cv2_oracle.extract(oracle_mod.lib(), img, 600, 1.2, 8, 20, 7, rects=rects, lap=(100, 250))
    assert rc == 0 and mono == mono2
    assert k.tobytes() == k2.tobytes()
    assert np.array_equal(d, d2)


def test_oracle_params_tum(oracle_mod):
    p = oracle_mod.Extractor(1000, 1.2, 8, 20, 7).params()
    assert p["quota"].tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert p["umax"].tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert oracle_mod.Extractor(1200, 1.2, 8, 20, 7).params()["quota"].tolist() == [261, 217, 181, 151, 126, 105, 87, 72]
    assert oracle_mod.Extractor(2000, 1.2, 8, 20, 7).params()["quota"].tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
    assert oracle_mod.Extractor(8000, 1.2, 8, 20, 7).params()["quota"].tolist() == [1737, 1448, 1207, 1005, 838, 698, 582, 485]


def test_oracle_edge_cases(oracle_mod):
    ex = oracle_mod.Extractor(1000, 1.2, 8, 20, 7)
    rc, k, d, mono = ex.extract(np.zeros((0, 0), np.uint8))
    assert rc == -1                                                               # empty image (:1129)
    rc, k, d, mono = ex.extract(np.full((240, 320), 128, np.uint8))
    assert rc == 0 and len(k) == 0 and mono == 0                                  # flat image → no keypoints
    # level sizes of the 640×480 pyramid
    from dani_slam_b200 import synth
    ex.extract(synth.throughput_frame(0))
    assert [ex.level_size(l) for l in range(8)] == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231), (257, 193), (214, 161), (179, 134)]


def test_knn_golden(oracle_mod):
    import os
    from conftest import GOLDEN
    from dani_slam_b200 import synth
    g = np.load(os.path.join(GOLDEN, "knn_cv2_s1234.npz"))
    q, db = synth.knn_case(int(g["nq"]), int(g["ndb"]), seed=int(g["seed"]), planted_frac=0.1)
    idx, dist = oracle_mod.knn2(q, db, nthreads=4)
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"])
    assert np.array_equal(oracle_mod.ratio_test(dist), g["keep"])
    assert g["keep"].sum() > 10
