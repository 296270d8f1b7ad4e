"""GPU parity tests of the extractor: CUDA path (through the C ABI) vs the CPU oracle and the golden
vectors — bit-exact for pyramid pixels, candidates, selected keypoints, final 28-byte keypoint records
(angles included: same fp32 formulation, so the 1e-3° tolerance of north_star is met with 0 difference),
descriptors and the mono index."""
import hashlib

import numpy as np
import pytest

from conftest import golden_cases, golden_frame, load_golden

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _xyz(k):
    return np.stack([k["x"], k["y"], k["response"]], axis=1).astype(np.float32)


@pytest.mark.parametrize("name", golden_cases())
def test_cuda_matches_golden_and_oracle_stage_by_stage(orbx_mod, oracle_mod, name):
    g = load_golden(name)
    img = golden_frame(g)
    nf, lap = int(g["nfeatures"]), tuple(int(v) for v in g["lap"])
    ex = orbx_mod.ORBextractor(nf, 1.2, 8, 20, 7, max_width=img.shape[1], max_height=img.shape[0], max_batch=1)
    ex.mvDynamicArea = [tuple(int(v) for v in r) for r in g["rects"]]
    mono, k, d = ex(img, None, lap)
    ref = oracle_mod.Extractor(nf, 1.2, 8, 20, 7)
    rc, rk, rd, rmono = ref.extract(img, rects=g["rects"], lap=lap, cap=nf + 200)
    for l in range(8):
        assert sha(ex.mvImagePyramid(l, padded=True)) == str(g["pyr_sha"][l]), f"pyramid level {l}"
        assert np.array_equal(ex.mvImagePyramid(l), ref.level(l))
        if str(g["blur_sha"][l]):
            assert sha(ex.blurred(l)) == str(g["blur_sha"][l]), f"blurred level {l}"
        assert sha(_xyz(ex.candidates(l))) == str(g["cand_sha"][l]), f"candidates level {l}"
        assert sha(_xyz(ex.selected(l))) == str(g["sel_sha"][l]), f"selected level {l}"
    assert mono == int(g["mono"]) == rmono
    assert k.tobytes() == g["kps"].tobytes() == rk.tobytes()
    assert np.array_equal(d, g["desc"]) and np.array_equal(d, rd)
    # angle tolerance stated by north_star (met exactly)
    assert np.max(np.abs(k["angle"] - rk["angle"]), initial=0) <= 1e-3


def test_batch_equals_per_frame_oracle(orbx_mod, oracle_mod):
    from dani_slam_b200 import synth
    B = 12
    imgs = np.stack([synth.parity_frame(100 + i) if i % 3 == 0 else synth.throughput_frame(100 + i) for i in range(B)])
    ex = orbx_mod.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=5)  # forces chunking 5+5+2
    n, mono, kps, desc = ex.extract_batch(imgs, (0, 0))
    ref = oracle_mod.Extractor(1000, 1.2, 8, 20, 7)
    seen = set()
    for b in range(B):
        rc, rk, rd, rmono = ref.extract(imgs[b])
        assert n[b] == len(rk) and mono[b] == rmono
        assert kps[b, : n[b]].tobytes() == rk.tobytes(), f"frame {b}"
        assert np.array_equal(desc[b, : n[b]], rd), f"frame {b}"
        seen.add(sha(rd))
    assert len(seen) == B  # the frames really are different


def test_device_resident_batch_api(orbx_mod, oracle_mod):
    import torch
    from dani_slam_b200 import synth
    B, H, W, cap = 6, 480, 752, 1400
    imgs = np.stack([synth.throughput_frame(200 + i, W, H) for i in range(B)])
    dev = torch.device("cuda", 0)
    d_img = torch.from_numpy(imgs).to(dev)
    d_k = torch.zeros((B, cap, 7), dtype=torch.float32, device=dev)
    d_d = torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev)
    d_n = torch.zeros(B, dtype=torch.int32, device=dev)
    d_m = torch.zeros(B, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ex = orbx_mod.ORBextractor(1200, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
    ex.extract_batch_device(d_img.data_ptr(), H * W, B, H, W, W, d_k.data_ptr(), d_d.data_ptr(), cap, d_n.data_ptr(), d_m.data_ptr())
    ex.sync()
    n, m = d_n.cpu().numpy(), d_m.cpu().numpy()
    k = d_k.cpu().numpy().view(np.uint8).reshape(B, cap, 28)
    dd = d_d.cpu().numpy()
    ref = oracle_mod.Extractor(1200, 1.2, 8, 20, 7)
    for b in range(B):
        rc, rk, rd, rmono = ref.extract(imgs[b])
        assert n[b] == len(rk) and m[b] == rmono
        assert k[b, : n[b]].tobytes() == rk.tobytes()
        assert np.array_equal(dd[b, : n[b]], rd)


def test_edge_cases_and_error_codes(orbx_mod, oracle_mod):
    from dani_slam_b200 import synth
    ex = orbx_mod.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=2)
    mono, k, d = ex(np.zeros((0, 0), np.uint8))
    assert mono == -1 and len(k) == 0                                    # empty image → -1 (:1129)
    mono, k, d = ex(np.full((240, 320), 77, np.uint8))
    assert mono == 0 and len(k) == 0 and d.shape == (0, 32)              # no keypoints → empty outputs (:1147)
    with pytest.raises(orbx_mod.OrbxError) as e:                          # pyramid level too small
        ex(np.zeros((60, 60), np.uint8))
    assert e.value.code == orbx_mod.ERR_GEOMETRY
    with pytest.raises(orbx_mod.OrbxError) as e:                          # larger than the handle's max size
        ex(np.zeros((481, 640), np.uint8))
    assert e.value.code == orbx_mod.ERR_ARG
    ex.cap = 100                                                          # capacity error reports the need
    with pytest.raises(orbx_mod.OrbxError) as e:
        ex(synth.throughput_frame(1))
    assert e.value.code == orbx_mod.ERR_CAPACITY
    ex.cap = 1200
    # the handle keeps working after errors and after a size change
    for (w, h, seed) in [(640, 480, 4), (401, 333, 5), (640, 480, 6)]:
        img = synth.parity_frame(seed, w, h)
        mono, k, d = ex(img, None, (0, 1000))
        rc, rk, rd, rmono = oracle_mod.Extractor(1000, 1.2, 8, 20, 7).extract(img, lap=(0, 1000))
        assert mono == rmono and k.tobytes() == rk.tobytes() and np.array_equal(d, rd)


def test_strided_input_and_other_parameters(orbx_mod, oracle_mod):
    from dani_slam_b200 import synth
    big = synth.throughput_frame(9, 800, 600)
    view = big[50:530, 80:720]                                            # non-contiguous rows (step 800)
    for (nf, sfac, nl, ini, mn) in [(1000, 1.2, 8, 20, 7), (500, 1.5, 4, 30, 10), (300, 1.1, 3, 12, 12), (2000, 1.2, 1, 20, 7)]:
        ex = orbx_mod.ORBextractor(nf, sfac, nl, ini, mn, max_width=640, max_height=480)
        mono, k, d = ex(view)
        rc, rk, rd, rmono = oracle_mod.Extractor(nf, sfac, nl, ini, mn).extract(np.ascontiguousarray(view), cap=nf + 300)
        assert rc == 0 and mono == rmono and k.tobytes() == rk.tobytes() and np.array_equal(d, rd), (nf, sfac, nl)
        p = oracle_mod.Extractor(nf, sfac, nl, ini, mn).params()
        assert np.array_equal(ex.GetScaleFactors(), p["sf"]) and np.array_equal(ex.GetInverseScaleFactors(), p["inv"])
        assert np.array_equal(ex.GetScaleSigmaSquares(), p["sigma2"]) and np.array_equal(ex.GetInverseScaleSigmaSquares(), p["inv_sigma2"])
        assert np.array_equal(ex.features_per_level(), p["quota"])


def test_dynamic_area_and_lapping_variants(orbx_mod, oracle_mod):
    from dani_slam_b200 import synth
    img = synth.throughput_frame(31)
    ex = orbx_mod.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480)
    ref = oracle_mod.Extractor(1000, 1.2, 8, 20, 7)
    cases = [([], (0, 0)), ([(0, 0, 640, 480)], (0, 0)), ([(200, 100, 300, 250)], (100, 400)),
             ([(10, 10, 5, 5), (600, 440, 40, 40), (320, 0, 1, 480)], (0, 1000)), ([(-50, -50, 100, 100)], (639, 639))]
    for rects, lap in cases:
        ex.mvDynamicArea = rects
        mono, k, d = ex(img, None, lap)
        rc, rk, rd, rmono = ref.extract(img, rects=rects, lap=lap)
        assert mono == rmono and k.tobytes() == rk.tobytes() and np.array_equal(d, rd), (rects, lap)
    ex.mvDynamicArea = [(0, 0, 640, 480)]
    mono, k, d = ex(img)
    assert len(k) == 0                                                    # everything deleted


def test_tie_heavy_and_regular_patterns(orbx_mod, oracle_mod):
    """Regular dot grids / checkerboards: equal FAST scores everywhere → NMS ties, quadtree (size, UL.x)
    ties resolved by libstdc++'s std::sort order, first-maximum ties in best-per-node."""
    H, W = 480, 640
    frames = []
    a = np.full((H, W), 90, np.uint8); a[5::11, 5::11] = 230; a[6::11, 5::11] = 230; frames.append(a)
    yy, xx = np.mgrid[0:H, 0:W]
    frames.append((((yy // 7 + xx // 7) & 1) * 140 + 50).astype(np.uint8))
    c = np.full((H, W), 128, np.uint8); c[::13, ::17] = 10; c[3::13, 5::17] = 250; frames.append(c)
    ex = orbx_mod.ORBextractor(1000, 1.2, 8, 20, 7, max_width=W, max_height=H)
    ref = oracle_mod.Extractor(1000, 1.2, 8, 20, 7)
    for i, f in enumerate(frames):
        mono, k, d = ex(f)
        rc, rk, rd, rmono = ref.extract(f)
        assert len(rk) > 100, i
        assert mono == rmono and k.tobytes() == rk.tobytes() and np.array_equal(d, rd), i


def test_fast_two_sided_pixels_and_dense_corner_fields(orbx_mod, oracle_mod):
    """The two-phase FAST kernel's rare paths: steep ramps and saddles make (almost) every pixel pass the compass
    pre-test on BOTH sides (its bounded two-sided queue then flushes every iteration); salt-and-pepper fields make
    most pixels corners (queues at their worst-case length); low-amplitude copies of them only fire at minThFAST."""
    H, W = 300, 420
    yy, xx = np.mgrid[0:H, 0:W]
    rng = np.random.default_rng(5)
    frames = [
        ((11 * xx + 9 * yy) % 256).astype(np.uint8),                                   # sawtooth ramp: two-sided everywhere
        ((13 * xx - 12 * yy) % 256).astype(np.uint8),
        (128 + 100 * np.sin(xx / 2.0) * np.sin(yy / 2.0)).astype(np.uint8),            # saddles between the bumps
        rng.choice(np.array([0, 255], np.uint8), size=(H, W)),                          # salt and pepper
        rng.integers(0, 256, (H, W), dtype=np.uint8),                                   # white noise
        (100 + ((11 * xx + 9 * yy) % 16)).astype(np.uint8),                            # amplitude 15: only the minTh pass
        (100 + rng.integers(0, 2, (H, W)) * 12).astype(np.uint8),                       # amplitude 12 salt and pepper
        np.where((xx // 3 + yy // 3) % 2 == 0, 40, 215).astype(np.uint8),               # 3-px checkerboard
    ]
    ex = orbx_mod.ORBextractor(1500, 1.2, 8, 20, 7, max_width=W, max_height=H)
    ref = oracle_mod.Extractor(1500, 1.2, 8, 20, 7)
    for i, f in enumerate(frames):
        mono, k, d = ex(f)
        rc, rk, rd, rmono = ref.extract(f, cap=2000)
        for l in range(8):
            assert sha(_xyz(ex.candidates(l))) == sha(_xyz(ref.candidates(l))), (i, l)
        assert mono == rmono and k.tobytes() == rk.tobytes() and np.array_equal(d, rd), i
    # and in one batch (different cells of a launch take different paths)
    exb = ex.__class__(1500, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=len(frames))
    n, mono, kps, desc = exb.extract_batch(np.stack(frames))
    for i, f in enumerate(frames):
        rc, rk, rd, rmono = ref.extract(f, cap=2000)
        assert n[i] == len(rk) and kps[i, : n[i]].tobytes() == rk.tobytes() and np.array_equal(desc[i, : n[i]], rd), i
    # cells with more candidates than the two-phase kernel's queue holds were redone by the single-phase kernel …
    import ctypes
    exb.L.orbx_debug_dense_count.argtypes = [ctypes.c_void_p]
    assert exb.L.orbx_debug_dense_count(exb.h) > 50
    # … and ordinary frames never take that path
    from dani_slam_b200 import synth
    exb.extract_batch(np.stack([synth.throughput_frame(s, W, H) for s in range(8)]))
    assert exb.L.orbx_debug_dense_count(exb.h) == 0


def test_full_size_4k_frame(orbx_mod, oracle_mod):
    """BASELINE config 5 size: one 3840×2160 frame, nFeatures=8000 — full parity plus structural properties."""
    from dani_slam_b200 import synth
    img = synth.throughput_frame(1, 3840, 2160)
    ex = orbx_mod.ORBextractor(8000, 1.2, 8, 20, 7, max_width=3840, max_height=2160)
    ex.cap = 8200
    mono, k, d = ex(img)
    assert 8000 <= len(k) <= 8000 + 16                                   # every level fills its quota (+≤2 each)
    assert np.all(np.diff(k["octave"]) >= 0)                             # level-major order when nothing laps
    assert (k["x"] >= 19).all() and (k["x"] <= 3840 - 19).all()
    rc, rk, rd, rmono = oracle_mod.Extractor(8000, 1.2, 8, 20, 7).extract(img, cap=8200)
    assert rc == 0 and mono == rmono and k.tobytes() == rk.tobytes() and np.array_equal(d, rd)


def test_clustered_corners_deep_quadtree_fallback(orbx_mod, oracle_mod):
    """Corners confined to tiny regions force the quadtree far deeper than the histogram table of the fast
    kernel (QT_DMAX levels), so the flagged (frame, level) pairs are redone by the general kernel; also a frame
    with a single textured blob, and one with two corners only."""
    from dani_slam_b200 import synth
    H, W = 480, 640
    rng = np.random.default_rng(77)
    frames = []
    a = np.full((H, W), 120, np.uint8)
    a[200:236, 300:340] = rng.integers(0, 256, (36, 40), dtype=np.uint8)          # one 40×36 noise blob
    frames.append(a)
    b = np.full((H, W), 90, np.uint8)
    for (y, x) in [(60, 70), (61, 400), (300, 90), (420, 600), (240, 320)]:
        b[y:y + 14, x:x + 14] = rng.integers(0, 256, (14, 14), dtype=np.uint8)    # five tiny blobs
    frames.append(b)
    c = np.full((H, W), 50, np.uint8)
    c[100:110, 100:110] = 250                                                      # a lone bright square
    c[300:305, 500:520] = 0
    frames.append(c)
    d = synth.throughput_frame(5)
    d[:, :] = np.where(np.add.outer(np.arange(H), np.arange(W)) % 97 < 90, 128, d)  # thin diagonal stripes of texture
    frames.append(np.ascontiguousarray(d))
    ex = orbx_mod.ORBextractor(1000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=4)
    ref = oracle_mod.Extractor(1000, 1.2, 8, 20, 7)
    n, mono, kps, desc = ex.extract_batch(np.stack(frames))
    for i, f in enumerate(frames):
        rc, rk, rd, rmono = ref.extract(f)
        assert n[i] == len(rk) and mono[i] == rmono, i
        assert kps[i, : n[i]].tobytes() == rk.tobytes() and np.array_equal(desc[i, : n[i]], rd), i
    assert n[0] >= 10 and n[2] >= 4
    ex.L.orbx_debug_deep_count.argtypes = [__import__('ctypes').c_void_p]
    assert ex.L.orbx_debug_deep_count(ex.h) > 0                                   # the fallback kernel really ran
    ex.extract_batch(np.stack([synth.throughput_frame(1), synth.throughput_frame(2)]))
    assert ex.L.orbx_debug_deep_count(ex.h) == 0                                  # dense frames stay on the fast path


def test_legacy_quadtree_kernel_gives_the_same_result(orbx_mod, oracle_mod, monkeypatch):
    """ORBX_LEGACY_QUADTREE=1 routes every (frame, level) through the general quadtree kernel."""
    from dani_slam_b200 import synth
    monkeypatch.setenv("ORBX_LEGACY_QUADTREE", "1")
    img = synth.parity_frame(61)
    ex = orbx_mod.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480)
    mono, k, d = ex(img, None, (0, 1000))
    rc, rk, rd, rmono = oracle_mod.Extractor(1000, 1.2, 8, 20, 7).extract(img, lap=(0, 1000))
    assert mono == rmono and k.tobytes() == rk.tobytes() and np.array_equal(d, rd)


def test_single_phase_fast_kernel_gives_the_same_result(orbx_mod, oracle_mod, monkeypatch):
    """ORBX_FAST_V1=1 selects the earlier FAST kernel that evaluates the exact measure at every pixel: an independent
    implementation of the same step, kept as a cross-check of the two-phase kernel."""
    from dani_slam_b200 import synth
    monkeypatch.setenv("ORBX_FAST_V1", "1")
    ref = oracle_mod.Extractor(1000, 1.2, 8, 20, 7)
    ex = orbx_mod.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480)
    for img in (synth.parity_frame(62), synth.throughput_frame(63)):
        mono, k, d = ex(img, None, (0, 0))
        rc, rk, rd, rmono = ref.extract(img)
        for l in range(8):
            assert sha(_xyz(ex.candidates(l))) == sha(_xyz(ref.candidates(l))), l
        assert mono == rmono and k.tobytes() == rk.tobytes() and np.array_equal(d, rd)


def test_left_right_extractors_run_concurrently_in_two_threads(orbx_mod, oracle_mod):
    """The reference extracts the left and right image in two std::threads with two extractor instances
    (src/Frame.cc:124-127); handles share no mutable state, so results must not depend on the interleaving."""
    import threading
    from dani_slam_b200 import synth
    left = [synth.throughput_frame(300 + i, 752, 480) for i in range(6)]
    right = [synth.stereo_right(f, 300 + i) for i, f in enumerate(left)]
    exL = orbx_mod.ORBextractor(1200, 1.2, 8, 20, 7, max_width=752, max_height=480)
    exR = orbx_mod.ORBextractor(1200, 1.2, 8, 20, 7, max_width=752, max_height=480)
    out = {"L": [], "R": []}

    def run(ex, frames, key):
        for _ in range(3):                                   # repeat to widen the overlap window
            res = [ex(f) for f in frames]
        out[key] = res

    tl = threading.Thread(target=run, args=(exL, left, "L"))
    tr = threading.Thread(target=run, args=(exR, right, "R"))
    tl.start(); tr.start(); tl.join(); tr.join()
    ref = oracle_mod.Extractor(1200, 1.2, 8, 20, 7)
    for frames, key in ((left, "L"), (right, "R")):
        for f, (mono, k, d) in zip(frames, out[key]):
            rc, rk, rd, rmono = ref.extract(f)
            assert mono == rmono and k.tobytes() == rk.tobytes() and np.array_equal(d, rd)


def test_other_pyramid_shapes(orbx_mod, oracle_mod):
    """Level counts, scale factors and thresholds away from the TUM1 defaults (incl. a single level, 12 levels,
    scale 2.0 which exercises the wide-source path of the pyramid kernel, and minTh > iniTh)."""
    from dani_slam_b200 import synth
    img = synth.parity_frame(71, 800, 600)
    for (nf, sfac, nl, ini, mn) in [(1500, 1.2, 12, 20, 7), (800, 2.0, 4, 20, 7), (600, 1.05, 6, 15, 5), (1000, 1.2, 8, 7, 20), (50, 1.3, 5, 40, 40),
                                    (0, 1.2, 8, 20, 7)]:
        ex = orbx_mod.ORBextractor(nf, sfac, nl, ini, mn, max_width=800, max_height=600)
        mono, k, d = ex(img)
        rc, rk, rd, rmono = oracle_mod.Extractor(nf, sfac, nl, ini, mn).extract(img, cap=nf + 400)
        assert rc == 0 and mono == rmono and k.tobytes() == rk.tobytes() and np.array_equal(d, rd), (nf, sfac, nl, ini, mn)


def test_randomised_geometries_and_parameters(orbx_mod, oracle_mod):
    """48 seeded draws of image size (down to the smallest a pyramid accepts), feature count, level count, scale factor,
    thresholds, lapping window and dynamic rectangles; content mixes noise, ramps, flats and dots.  Small and odd sizes
    produce one-cell levels, clipped cells narrower than the FAST ring and cells of very different heights in one launch."""
    from dani_slam_b200 import synth
    rng = np.random.default_rng(2024)
    done = 0
    for case in range(48):
        nl = int(rng.integers(1, 9))
        sfac = float(rng.choice([1.1, 1.2, 1.25, 1.4, 1.7]))
        w = int(rng.integers(70, 420)); h = int(rng.integers(70, 330))
        nf = int(rng.choice([40, 200, 700, 1500]))
        ini = int(rng.integers(8, 40)); mn = int(rng.integers(3, ini + 1))
        kind = case % 4
        if kind == 0:
            img = synth.parity_frame(case, w, h)
        elif kind == 1:
            img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        elif kind == 2:
            yy, xx = np.mgrid[0:h, 0:w]
            img = ((int(rng.integers(3, 15)) * xx + int(rng.integers(3, 15)) * yy) % 256).astype(np.uint8)
        else:
            img = synth.throughput_frame(case, w, h)
        rects = [(int(rng.integers(0, w)), int(rng.integers(0, h)), int(rng.integers(5, 90)), int(rng.integers(5, 90))) for _ in range(int(rng.integers(0, 3)))]
        lap = [(0, 0), (0, 1000), (w // 3, 2 * w // 3)][case % 3]
        try:
            ex = orbx_mod.ORBextractor(nf, sfac, nl, ini, mn, max_width=w, max_height=h)
            ex.mvDynamicArea = rects
            mono, k, d = ex(img, None, lap)
        except orbx_mod.OrbxError as e:
            assert e.code == orbx_mod.ERR_GEOMETRY or "orbx_create: image too" in str(e), (case, e)   # geometries the reference itself faults on
            continue
        rc, rk, rd, rmono = oracle_mod.Extractor(nf, sfac, nl, ini, mn).extract(img, rects=rects, lap=lap, cap=nf + 400)
        assert rc == 0 and mono == rmono and k.tobytes() == rk.tobytes() and np.array_equal(d, rd), (case, w, h, nf, sfac, nl, ini, mn)
        done += 1
    assert done >= 30


def test_many_seeded_frames_statistical_parity(orbx_mod, oracle_mod):
    """256 seeded frames (throughput + parity generators, three lapping windows): every keypoint record and
    descriptor equals the oracle's.  Rare paths (rounding of rotated taps at .5, std::sort tie permutations at the
    quota cut, drift on split lines) only show up at this scale."""
    from concurrent.futures import ThreadPoolExecutor
    from dani_slam_b200 import synth
    B = 256
    imgs = np.stack([synth.parity_frame(1000 + i) if i % 4 == 3 else synth.throughput_frame(1000 + i) for i in range(B)])
    ex = orbx_mod.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=B)
    for lap in [(0, 0), (0, 1000), (250, 400)]:
        n, mono, kps, desc = ex.extract_batch(imgs, lap)

        def check(b):
            rc, rk, rd, rmono = oracle_mod.Extractor(1000, 1.2, 8, 20, 7).extract(imgs[b], lap=lap)
            return (n[b] == len(rk) and mono[b] == rmono and kps[b, : n[b]].tobytes() == rk.tobytes()
                    and np.array_equal(desc[b, : n[b]], rd))

        step = 1 if lap == (0, 0) else 4              # full set once, a quarter for the other lapping windows
        with ThreadPoolExecutor(16) as pool:
            ok = list(pool.map(check, range(0, B, step)))
        assert all(ok), (lap, [i * step for i, v in enumerate(ok) if not v][:10])


def test_single_frame_path_graph_replay_and_eager_agree(orbx_mod, oracle_mod, monkeypatch):
    """orbx_extract replays copy-in + kernels + copy-out as one CUDA graph from the third call with the same geometry: eight calls on
    alternating frames, changing lapping window, dynamic rectangles and image size (graph re-capture) all equal the oracle, and the
    eager form (ORBX_NO_GRAPH) gives the same bytes."""
    from dani_slam_b200 import synth
    frames = [synth.parity_frame(60 + i, 640, 480) for i in range(3)]
    small = synth.parity_frame(70, 500, 375)
    ref = oracle_mod.Extractor(1000, 1.2, 8, 20, 7)

    def run(ex):
        out = []
        for k in range(8):
            img = frames[k % 3]
            lap = (0, 0) if k < 5 else (200, 400)
            ex.mvDynamicArea = [(100, 80, 150, 120)] if k == 6 else []
            out.append((ex(img, None, lap), img, list(ex.mvDynamicArea), lap))
        out.append((ex(small, None, (0, 0)), small, [], (0, 0)))            # smaller image: new geometry, new graph
        for k in range(4):
            out.append((ex(small, None, (0, 0)), small, [], (0, 0)))
        return out

    got = run(orbx_mod.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480))
    for (mono, k, d), img, rects, lap in got:
        rc, rk, rd, rmono = ref.extract(img, rects=rects, lap=lap)
        assert rc == 0 and mono == rmono and k.tobytes() == rk.tobytes() and np.array_equal(d, rd)
    monkeypatch.setenv("ORBX_NO_GRAPH", "1")
    eager = run(orbx_mod.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480))
    for ((m1, k1, d1), *_), ((m2, k2, d2), *_) in zip(got, eager):
        assert m1 == m2 and k1.tobytes() == k2.tobytes() and np.array_equal(d1, d2)
