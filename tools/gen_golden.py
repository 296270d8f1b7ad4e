#!/usr/bin/env python
"""Generate tests/golden/*.npz — golden vectors for the ORB front-end hot path.

The reference ships no fixtures for this path, so the vectors are produced here from the
REAL OpenCV primitives of the container's cv2 wheel driven by the reference's orchestration as restated
in oracle/cv2_oracle.py (which needs only libstdc++'s std::sort from the C++ oracle for the quadtree's
tie order).  Run in the build container:  python tools/gen_golden.py [case ...]
Each .npz holds, for one seeded synthetic frame: the frame parameters, SHA-256 digests of every pyramid
level (padded), blurred level, per-level candidate and selected lists, and the final keypoints (28-byte
records) + descriptors in full.  Matching cases hold cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) results.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dani_slam_b200 import synth  # noqa: E402
from oracle import cv2_oracle, oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = [
    # name, generator, seed, W, H, nfeatures, rects, lap
    ("tum_parity_s3", "parity", 3, 640, 480, 1000, [], (0, 0)),
    ("tum_mono_s0", "throughput", 0, 640, 480, 1000, [], (0, 1000)),
    ("odd_rects_s5", "parity", 5, 641, 479, 1000, [(100, 80, 120, 90), (300, 200, 50, 50), (10, 400, 600, 30)], (0, 0)),
    ("kitti_s1", "throughput", 1, 1241, 376, 2000, [], (0, 0)),
    ("euroc_lap_s2", "throughput", 2, 752, 480, 1200, [], (200, 500)),
    ("small_s9", "parity", 9, 320, 240, 500, [], (0, 0)),
    ("uhd_s4", "throughput", 4, 3840, 2160, 8000, [], (0, 0)),        # BASELINE config 5 at full size
]


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def frame_for(kind, seed, W, H):
    return synth.parity_frame(seed, W, H) if kind == "parity" else synth.throughput_frame(seed, W, H)


def main():
    import cv2
    os.makedirs(OUT, exist_ok=True)
    L = oracle.lib()
    only = set(sys.argv[1:])                      # optional: regenerate just the named cases
    for name, kind, seed, W, H, nf, rects, lap in CASES:
        if only and name not in only:
            continue
        img = frame_for(kind, seed, W, H)
        taps = {}
        kps, desc, mono = cv2_oracle.extract(L, img, nf, 1.2, 8, 20, 7, rects=rects, lap=lap, taps=taps)
        d = dict(kind=kind, seed=seed, W=W, H=H, nfeatures=nf, rects=np.asarray(rects, np.int32).reshape(-1, 4),
                 lap=np.asarray(lap, np.int32), frame_sha=sha(img), kps=kps, desc=desc, mono=mono,
                 cv2_version=cv2.__version__)
        pyr, blur, cand, sel = [], [], [], []
        for l in range(8):
            pyr.append(sha(taps["planes"][l]))
            b = taps["blur"][l] if l < len(taps["blur"]) else None
            blur.append(sha(b) if b is not None else "")
            X, Y, R = taps["cand"][l]
            cand.append(sha(np.stack([X, Y, R], axis=1).astype(np.float32)))
            s = taps["sel"][l]
            sel.append(sha(np.stack([s["x"], s["y"], s["response"]], axis=1).astype(np.float32)))
        d.update(pyr_sha=np.array(pyr), blur_sha=np.array(blur), cand_sha=np.array(cand), sel_sha=np.array(sel))
        np.savez_compressed(os.path.join(OUT, f"extract_{name}.npz"), **d)
        print(name, len(kps), mono)
    if only and "knn" not in only:
        return
    # matching: real cv2 BFMatcher on a planted/tie-heavy case
    q, db = synth.knn_case(300, 20000, seed=1234, planted_frac=0.1)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    m = bf.knnMatch(q, db, k=2)
    idx = np.array([[mm[0].trainIdx, mm[1].trainIdx] for mm in m], np.int32)
    dist = np.array([[int(mm[0].distance), int(mm[1].distance)] for mm in m], np.int32)
    keep = np.array([mm[0].distance < mm[1].distance * 0.7 for mm in m], bool)
    np.savez_compressed(os.path.join(OUT, "knn_cv2_s1234.npz"), nq=300, ndb=20000, seed=1234, idx=idx, dist=dist, keep=keep,
                        cv2_version=cv2.__version__)
    print("knn", idx.shape, keep.sum())


if __name__ == "__main__":
    main()
