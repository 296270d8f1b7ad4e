"""Generates tests/golden/match_ref.npz — outputs of the REFERENCE's own matcher / Frame-grid code
(oracle/_ref/libref_match.so = src/ORBmatcher.cc and src/Frame.cc line ranges compiled unmodified, see oracle/Makefile)
on seeded synthetic scenes (dani_slam_b200/synth.py).  Run in the build container, where /root/reference exists:

    python tools/gen_golden_match.py

The fixture lets the oracle and the CUDA path be checked against reference OUTPUTS where oracle/_ref is absent.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dani_slam_b200 import synth  # noqa: E402
from oracle import ref_binding as R  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
from match_cases import AREA_CASES, BOUNDS, BOW_CASES, INIT_CASES, SBP_CASES, area_queries, histo_cases, sha, tail_case  # noqa: E402


def main():
    out = {}
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (200, 32), dtype=np.uint8); b = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    b[:5] = a[:5]; b[5] = ~a[5]
    out["dd_a"], out["dd_b"] = a, b
    out["dd"] = np.array([R.descriptor_distance(a[i], b[i]) for i in range(200)], np.int32)
    c = R.match_constants()
    out["constants"] = np.array([c["TH_LOW"], c["TH_HIGH"], c["HISTO_LENGTH"]], np.int32)
    h = histo_cases()
    out["histo"] = h
    out["maxima"] = np.array([R.three_maxima(x) for x in h], np.int32)
    for i, (n, nq, seed, lv) in enumerate(AREA_CASES):
        k = synth.keypoint_records(n, seed)
        off, cand = R.features_in_area(k, BOUNDS, area_queries(nq, seed), *lv)
        out[f"area{i}_in"] = np.array(sha(k), dtype="U64"); out[f"area{i}_off"] = off; out[f"area{i}_cand"] = cand
    for i, (n1, n2, seed, ratio, ori, win) in enumerate(INIT_CASES):
        k1, d1, k2, d2 = synth.init_scene(n1, n2, seed)
        n, m12, prev = R.search_for_initialization(k1, d1, k2, d2, BOUNDS, np.stack([k1["x"], k1["y"]], 1), win, ratio, ori)
        out[f"init{i}_in"] = np.array(sha(k1, d1, k2, d2), dtype="U64"); out[f"init{i}_n"] = np.int32(n); out[f"init{i}_m12"] = m12
        out[f"init{i}_prev"] = prev
    for i, (n, m, seed, st, th) in enumerate(SBP_CASES):
        s = synth.projection_scene(n, m, seed, stereo=st)
        nm, asg = R.search_by_projection(s["kps"], s["desc"], BOUNDS, s["scale_factors"], s["mp_proj5"], s["mp_level"], s["mp_flags"], s["mp_obs"],
                                         s["mp_desc"], 0.8, th, True, 50.0, s["u_right"], s["kp_obs"])
        out[f"sbp{i}_in"] = np.array(sha(s["kps"], s["desc"], s["mp_proj5"], s["mp_desc"]), dtype="U64")
        out[f"sbp{i}_n"] = np.int32(nm); out[f"sbp{i}_assigned"] = asg
    for i, (nk, nf, seed, ratio, ori) in enumerate(BOW_CASES):
        s = synth.bow_scene(nk, nf, seed)
        nm, asg = R.search_by_bow(s["kf_kps"], s["kf_desc"], s["kf_mp"], s["kf_fv"], s["f_kps"], s["f_desc"], s["f_fv"], ratio, ori)
        out[f"bow{i}_in"] = np.array(sha(s["kf_kps"], s["kf_desc"], s["kf_mp"], *s["kf_fv"], s["f_kps"], s["f_desc"], *s["f_fv"]), dtype="U64")
        out[f"bow{i}_n"] = np.int32(nm); out[f"bow{i}_assigned"] = asg
    for i, (nk, nf, seed, ratio, ori) in enumerate(BOW_CASES):
        s = synth.bow_scene(nk, nf, seed)
        nm, m12 = R.search_by_bow_kf(s["kf_kps"], s["kf_desc"], s["kf_mp"], s["kf_fv"], s["f_kps"], s["f_desc"], s["f_mp"], s["f_fv"], ratio, ori)
        out[f"bowkf{i}_in"] = np.array(sha(s["kf_kps"], s["kf_desc"], s["kf_mp"], *s["kf_fv"], s["f_kps"], s["f_desc"], s["f_mp"], *s["f_fv"]), dtype="U64")
        out[f"bowkf{i}_n"] = np.int32(nm); out[f"bowkf{i}_m12"] = m12
    for i, seed in enumerate([1, 2]):
        uL, uR, iL, iR, dist = tail_case(seed)
        # the reference's matcher in this slot reports a similarity (distance = 1 - match.distance, src/Frame.cc:891); feeding
        # match.distance = 1 - d makes the reference's `distance` the Hamming distance d exactly (both subtractions are exact in fp32)
        n, ur, dp = R.stereo_tail(uL, uR, iL, iR, (np.float32(1.0) - dist.astype(np.float32)), 386.1448, 0.53716)
        out[f"tail{i}_n"] = np.int32(n); out[f"tail{i}_ur"] = ur; out[f"tail{i}_depth"] = dp
    path = os.path.join(ROOT, "tests", "golden", "match_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
