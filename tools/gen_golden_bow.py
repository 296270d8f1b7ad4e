#!/usr/bin/env python
"""Generate tests/golden/bow_*.npz and tests/golden/undistort_*.npz.

bow_*:       outputs of the REFERENCE'S OWN DBoW2 (Thirdparty/DBoW2 compiled unmodified into oracle/_ref/libref_bow.so,
             `make -C oracle ref`) for synthetic vocabularies written in the reference's text format and loaded with its
             own loadFromTextFile: per-descriptor word ids, BowVector (ids, double values), FeatureVector (CSR), for
             levelsup 4 as Frame::ComputeBoW uses and one other level.  The vocabulary and the query descriptors are
             regenerated from seeds by dani_slam_b200.synth; their SHA-256 is stored so that generator drift is caught.
undistort_*: cv2.undistortPoints (the container's cv2 wheel) on seeded points for the camera models of
             tests/test_undistort.py, plus the ComputeImageBounds corners.
Run in the build container (needs /root/reference for the DBoW2 build and cv2):  python tools/gen_golden_bow.py
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dani_slam_b200 import synth  # noqa: E402
from oracle import ref_binding  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

BOW_CASES = [
    ("k10_L3", dict(k=10, L=3, seed=101), 1200, 4),
    ("k10_L4_stop", dict(k=10, L=4, seed=102, stop_frac=0.2), 900, 4),
    ("k5_L6_ties", dict(k=5, L=6, seed=103, flips=6), 700, 4),
    ("k10_L3_l2", dict(k=10, L=3, seed=104, scoring=1), 500, 2),
    ("k10_L3_idf", dict(k=10, L=3, seed=105, weighting=2), 500, 1),
]

CAMS = [
    ("tum1", (517.306408, 516.469215, 318.643040, 255.313989), [0.262383, -0.953104, -0.005358, 0.002628, 1.163314], (640, 480)),
    ("euroc", (458.654, 457.296, 367.215, 248.375), [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], (752, 480)),
    ("rational", (535.4, 539.2, 320.1, 247.6), [0.05, -0.1, 0.002, -0.001, 0.02, 0.01, -0.02, 0.005], (640, 480)),
]


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    if not ref_binding.bow_available():
        raise SystemExit("oracle/_ref/libref_bow.so missing: run `make -C oracle ref` where /root/reference exists")
    for name, vargs, nq, levelsup in BOW_CASES:
        voc = synth.vocabulary(**vargs)
        q = synth.vocabulary_queries(voc, nq, seed=vargs["seed"] + 1)
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "voc.txt")
            synth.write_vocabulary_text(voc, path)
            r = ref_binding.Vocabulary(path).transform(q, levelsup)
        np.savez_compressed(os.path.join(OUT, f"bow_{name}.npz"), vargs=np.array(repr(vargs)), nq=nq, levelsup=levelsup,
                            voc_sha=sha(np.concatenate([voc["parent"].view(np.uint8), voc["is_leaf"], voc["desc"].ravel(), voc["weight"].view(np.uint8)])),
                            q_sha=sha(q), **r)
        print(name, "words", len(r["bow_ids"]), "nodes", len(r["fv_nodes"]))
    import cv2
    for name, (fx, fy, cx, cy), D, (w, h) in CAMS:
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
        Df = np.asarray(D, np.float32)
        rng = np.random.default_rng(77)
        pts = np.stack([rng.uniform(-40, w + 40, 5000), rng.uniform(-40, h + 40, 5000)], 1).astype(np.float32)
        corners = np.array([[0, 0], [w, 0], [0, h], [w, h]], np.float32)
        und = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, Df, None, K).reshape(-1, 2)
        uc = cv2.undistortPoints(corners.reshape(-1, 1, 2), K, Df, None, K).reshape(-1, 2)
        bounds = np.array([min(uc[0, 0], uc[2, 0]), max(uc[1, 0], uc[3, 0]), min(uc[0, 1], uc[1, 1]), max(uc[2, 1], uc[3, 1])], np.float32)
        np.savez_compressed(os.path.join(OUT, f"undistort_{name}.npz"), K=np.array([fx, fy, cx, cy], np.float32), D=Df, size=np.array([w, h]),
                            pts=pts, undistorted=und, bounds=bounds)
        print(name, "undistort ok")


if __name__ == "__main__":
    main()
