"""Pin the oracle's primitive restatements to the real OpenCV (cv2 wheel of this image): SURVEY.md App. A."""
import ctypes as C

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("sw,sh,dw,dh", [(640, 480, 533, 400), (533, 400, 444, 333), (179, 134, 149, 112),
                                         (1241, 376, 1034, 313), (101, 77, 84, 64), (64, 48, 97, 71), (50, 50, 50, 50),
                                         (333, 200, 111, 67)])
def test_resize_linear_matches_cv2(oracle_mod, sw, sh, dw, dh):
    rng = np.random.default_rng(sw * 7 + dh)
    src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
    ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
    out = np.zeros((dh, dw), np.uint8)
    oracle_mod.lib().orc_resize_linear_u8(_p(src), sw, sh, src.strides[0], _p(out), dw, dh, out.strides[0])
    assert np.array_equal(out, ref)


def test_resize_strided_source(oracle_mod):
    rng = np.random.default_rng(1)
    big = rng.integers(0, 256, (300, 400), dtype=np.uint8)
    src = big[19:-19, 19:-19]  # a ROI view like mvImagePyramid[level-1]
    dw, dh = round(src.shape[1] / 1.2), round(src.shape[0] / 1.2)
    ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
    out = np.zeros((dh, dw), np.uint8)
    oracle_mod.lib().orc_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(out), dw, dh, out.strides[0])
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("w,h", [(64, 48), (21, 33), (100, 20)])
def test_border_reflect101_matches_cv2(oracle_mod, w, h):
    rng = np.random.default_rng(w)
    src = rng.integers(0, 256, (h, w), dtype=np.uint8)
    ref = cv2.copyMakeBorder(src, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
    out = np.zeros_like(ref)
    oracle_mod.lib().orc_border_reflect101_u8(_p(src), w, h, src.strides[0], _p(out), out.strides[0], 19)
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("w,h,smooth", [(640, 480, False), (179, 134, True), (37, 41, False), (8, 9, False), (1241, 376, True)])
def test_gaussian7_matches_cv2(oracle_mod, w, h, smooth):
    rng = np.random.default_rng(w + h)
    src = rng.integers(0, 256, (h, w), dtype=np.uint8)
    if smooth:
        src = cv2.blur(src, (5, 5))
    ref = cv2.GaussianBlur(src.copy(), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    out = np.zeros_like(src)
    oracle_mod.lib().orc_gaussian7_u8(_p(src), w, h, src.strides[0], _p(out), out.strides[0])
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("th", [7, 20, 0, 40])
@pytest.mark.parametrize("w,h", [(46, 44), (41, 57), (7, 7), (8, 30), (120, 90)])
def test_fast9_nms_matches_cv2(oracle_mod, w, h, th):
    from dani_slam_b200 import synth
    frame = synth.parity_frame(w * 100 + h + th, 320, 240)
    det = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    for (x0, y0) in [(16, 16), (150, 30), (20, 130), (170, 140), (100, 100)]:
        roi = frame[y0:y0 + h, x0:x0 + w]
        kps = det.detect(roi)
        ref = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in kps], np.int32).reshape(-1, 3)
        out = np.zeros((4096, 3), np.int32)
        n = oracle_mod.lib().orc_fast9_nms(_p(roi), w, h, roi.strides[0], th, _p(out), 4096)
        assert n == len(ref)
        assert np.array_equal(out[:n], ref)


def test_fast_atan2_matches_cv2(oracle_mod):
    rng = np.random.default_rng(0)
    ys = rng.integers(-1_300_000, 1_300_000, 20000)
    xs = rng.integers(-1_300_000, 1_300_000, 20000)
    special = [(0, 0), (0, 5), (5, 0), (-5, 0), (0, -5), (-1, 1000000), (1, -1000000), (-1, -1), (7, 7), (-7, 7)]
    L = oracle_mod.lib()
    for y, x in list(zip(ys.tolist(), xs.tolist()))[:5000] + special:
        a = L.orc_fast_atan2(float(np.float32(y)), float(np.float32(x)))
        b = cv2.fastAtan2(float(np.float32(y)), float(np.float32(x)))
        assert np.float32(a).tobytes() == np.float32(b).tobytes(), (y, x, a, b)


def test_cvround_half_to_even(oracle_mod):
    L = oracle_mod.lib()
    for v, want in [(0.5, 0), (1.5, 2), (2.5, 2), (-0.5, 0), (-1.5, -2), (2.4999, 2), (1e6 + 0.5, 1000000)]:
        assert L.orc_cvround(v) == want


def test_knn2_matches_cv2_bfmatcher(oracle_mod):
    from dani_slam_b200 import synth
    q, db = synth.knn_case(200, 5000, seed=77, planted_frac=0.1)
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, db, k=2)
    ridx = np.array([[a.trainIdx, b.trainIdx] for a, b in m], np.int32)
    rdist = np.array([[int(a.distance), int(b.distance)] for a, b in m], np.int32)
    idx, dist = oracle_mod.knn2(q, db, nthreads=2)
    assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist)
    keep = oracle_mod.ratio_test(dist)
    assert np.array_equal(keep, np.array([a.distance < b.distance * 0.7 for a, b in m]))


def test_descriptor_distance_bit_hack(oracle_mod):
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (100, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (100, 32), dtype=np.uint8)
    L = oracle_mod.lib()
    for i in range(100):
        assert L.orc_descriptor_distance(_p(a[i]), _p(b[i])) == int(np.unpackbits(a[i] ^ b[i]).sum())
