"""Seeded matcher cases shared by tools/gen_golden_match.py (which stores the REFERENCE's outputs for them in
tests/golden/match_ref.npz) and the tests that check the oracle / the CUDA path against those outputs."""
import hashlib

import numpy as np

BOUNDS = (0.0, 0.0, 640.0, 480.0)
INIT_CASES = [(500, 300, 1, 0.9, True, 100), (2000, 2500, 2, 0.9, True, 100), (800, 50, 3, 0.6, False, 100), (1500, 1500, 6, 0.9, True, 30)]
SBP_CASES = [(2000, 1500, 1, False, 1.0), (1000, 3000, 2, True, 3.0), (50, 20, 3, False, 3.0), (3000, 3000, 5, True, 1.0)]
BOW_CASES = [(1000, 1200, 1, 0.7, True), (2000, 2000, 2, 0.75, True), (800, 900, 3, 0.9, False), (40, 40, 6, 0.6, True)]   # n_kf, n_f, seed, ratio, checkOri
AREA_CASES = [(3000, 500, 1, (-1, -1)), (3000, 500, 2, (0, 0)), (800, 300, 3, (1, 4)), (2000, 400, 4, (-1, 2))]


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def area_queries(nq, seed):
    rng = np.random.default_rng(seed)
    return np.stack([rng.uniform(-20, 660, nq), rng.uniform(-20, 500, nq), rng.choice([0.5, 3, 15, 100], nq)], 1).astype(np.float32)


def histo_cases():
    rng = np.random.default_rng(9)
    cases = [np.zeros(30, np.int32), np.r_[10, 1, 1, np.zeros(27)].astype(np.int32), np.r_[10, 0.99, 1, np.zeros(27)].astype(np.int32),
             np.r_[100, 10, 9, np.zeros(27)].astype(np.int32), np.r_[100, 9, 10, 11, np.zeros(26)].astype(np.int32),
             np.full(30, 5, np.int32), np.r_[np.zeros(27), 3, 3, 3].astype(np.int32)]
    cases += [rng.integers(0, 40, 30).astype(np.int32) for _ in range(40)]
    cases += [(rng.integers(0, 3, 30) * rng.integers(0, 200, 30)).astype(np.int32) for _ in range(40)]
    return np.stack(cases)


def tail_case(seed, n_left=1500, n_right=1400):
    rng = np.random.default_rng(seed)
    uL = rng.uniform(0, 1241, n_left).astype(np.float32)
    uR = rng.uniform(0, 1241, n_right).astype(np.float32)
    iL = rng.permutation(n_left)[: n_left * 2 // 3].astype(np.int32)
    iR = rng.integers(0, n_right, len(iL)).astype(np.int32)
    good = rng.random(len(iL)) < 0.7
    uR[iR[good]] = uL[iL[good]] - rng.choice([0.0, 0.0, 5.0, 30.0, 200.0, 800.0], good.sum()).astype(np.float32)
    dist = rng.integers(0, 120, len(iL)).astype(np.int32)
    return uL, uR, iL, iR, dist
