"""Frame::UndistortKeyPoints / ComputeImageBounds (src/Frame.cc:749-811): the oracle against the
real cv2.undistortPoints (CPU), and the CUDA path through the C ABI against the oracle (GPU).  Bit-exact floats."""
import numpy as np
import pytest

from oracle import oracle

# TUM1 / EuRoC / KITTI-like intrinsics and distortion sets (Examples/*/ *.yaml of the reference), plus a 4-coefficient set
CAMS = [
    ((517.306408, 516.469215, 318.643040, 255.313989), [0.262383, -0.953104, -0.005358, 0.002628, 1.163314], (640, 480)),
    ((458.654, 457.296, 367.215, 248.375), [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], (752, 480)),
    ((718.856, 718.856, 607.1928, 185.2157), [0.0, 0.0, 0.0, 0.0], (1241, 376)),          # D[0] == 0: copy
    ((535.4, 539.2, 320.1, 247.6), [0.05, -0.1, 0.002, -0.001, 0.02, 0.01, -0.02, 0.005], (640, 480)),   # rational model
]


def _points(n, size, seed):
    rng = np.random.default_rng(seed)
    w, h = size
    return np.stack([rng.uniform(-40, w + 40, n), rng.uniform(-40, h + 40, n)], 1).astype(np.float32)


@pytest.mark.parametrize("cam", CAMS, ids=lambda c: "fx%d_%dcoef" % (c[0][0], len(c[1])))
def test_oracle_matches_cv2(cam):
    cv2 = pytest.importorskip("cv2")
    (fx, fy, cx, cy), D, size = cam
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
    Df = np.asarray(D, np.float32)
    pts = _points(100000, size, 3)
    mine = oracle.undistort_points(pts, *K[[0, 1, 0, 1], [0, 1, 2, 2]], Df)
    ref = pts if Df[0] == 0 else cv2.undistortPoints(pts.reshape(-1, 1, 2), K, Df, None, K).reshape(-1, 2)
    assert np.array_equal(ref.view(np.uint32), mine.view(np.uint32))
    b = oracle.image_bounds(size[0], size[1], *K[[0, 1, 0, 1], [0, 1, 2, 2]], Df)
    if Df[0] == 0:
        assert list(b) == [0, size[0], 0, size[1]]
    else:
        c = np.array([[0, 0], [size[0], 0], [0, size[1]], [size[0], size[1]]], np.float32)
        u = cv2.undistortPoints(c.reshape(-1, 1, 2), K, Df, None, K).reshape(-1, 2)
        assert list(b) == [min(u[0, 0], u[2, 0]), max(u[1, 0], u[3, 0]), min(u[0, 1], u[1, 1]), max(u[2, 1], u[3, 1])]


@pytest.mark.gpu
@pytest.mark.parametrize("cam", CAMS, ids=lambda c: "fx%d_%dcoef" % (c[0][0], len(c[1])))
def test_cuda_matches_oracle(orbx_mod, cam):
    import torch
    (fx, fy, cx, cy), D, size = cam
    Kf = tuple(np.float32(v) for v in (fx, fy, cx, cy))
    Df = np.asarray(D, np.float32)
    m = orbx_mod.ORBmatcher()
    for n in (0, 1, 1003, 50000):
        kps = np.zeros(n, orbx_mod.KP_DTYPE)
        pts = _points(n, size, n)
        kps["x"], kps["y"] = pts[:, 0], pts[:, 1]
        kps["size"], kps["angle"], kps["response"], kps["octave"], kps["class_id"] = 31, 12.5, 40, 1, -1
        un = m.UndistortKeyPoints(kps, Kf, Df)
        ref = oracle.undistort_points(pts, *Kf, Df)
        assert np.array_equal(np.stack([un["x"], un["y"]], 1).view(np.uint32).reshape(-1, 2), ref.view(np.uint32).reshape(-1, 2))
        for f in ("size", "angle", "response", "octave", "class_id"):
            assert np.array_equal(un[f], kps[f])
    assert np.array_equal(m.ComputeImageBounds(size[0], size[1], Kf, Df), oracle.image_bounds(size[0], size[1], *Kf, Df))
    # device form on packed pairs, out of place
    pts = _points(4096, size, 9)
    d_in = torch.from_numpy(pts).cuda(); d_out = torch.zeros_like(d_in)
    m.undistort_points_device(d_in.data_ptr(), 2, len(pts), Kf, Df, d_out.data_ptr(), 2)
    m.sync()
    assert np.array_equal(d_out.cpu().numpy().view(np.uint32), oracle.undistort_points(pts, *Kf, Df).view(np.uint32))


import glob
import os

UND_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "undistort_*.npz")))


@pytest.mark.parametrize("path", UND_GOLDEN, ids=lambda p: os.path.basename(p)[10:-4])
def test_oracle_matches_golden_cv2_vectors(path):
    g = np.load(path)
    K, D = g["K"], g["D"]
    assert np.array_equal(oracle.undistort_points(g["pts"], *K, D).view(np.uint32), g["undistorted"].view(np.uint32))
    assert np.array_equal(oracle.image_bounds(int(g["size"][0]), int(g["size"][1]), *K, D), g["bounds"])


@pytest.mark.gpu
def test_cuda_matches_golden_cv2_vectors(orbx_mod):
    assert len(UND_GOLDEN) >= 3
    m = orbx_mod.ORBmatcher()
    for path in UND_GOLDEN:
        g = np.load(path)
        K, D, pts = tuple(g["K"]), g["D"], g["pts"]
        kps = np.zeros(len(pts), orbx_mod.KP_DTYPE)
        kps["x"], kps["y"] = pts[:, 0], pts[:, 1]
        un = m.UndistortKeyPoints(kps, K, D)
        assert np.array_equal(np.stack([un["x"], un["y"]], 1).view(np.uint32), g["undistorted"].view(np.uint32)), path
        assert np.array_equal(m.ComputeImageBounds(int(g["size"][0]), int(g["size"][1]), K, D), g["bounds"]), path
