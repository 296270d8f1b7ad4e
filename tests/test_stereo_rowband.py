"""Classical rectified-stereo association (slot Frame::ComputeStereoMatches, src/Frame.cc:813-915).

PARITY UNPINNED by definition: this tree fills the slot with LightGlue, so there is no reference code to run.  The oracle restates the
published upstream algorithm; here it is cross-checked against an independent numpy restatement (CPU) and the CUDA path must equal the
oracle bit for bit (GPU)."""
import numpy as np
import pytest

FX, B = 718.856, 0.53716          # Examples/Stereo/KITTI00-02.yaml: Camera1.fx, Stereo.b


def _pair(seed, W, H, nf, oracle_mod):
    from dani_slam_b200 import synth
    left = synth.throughput_frame(seed, W, H)
    right = synth.stereo_right(left, seed)
    eL, eR = oracle_mod.Extractor(nf, 1.2, 8, 20, 7), oracle_mod.Extractor(nf, 1.2, 8, 20, 7)
    _, kl, dl, _ = eL.extract(left, cap=nf + 200)
    _, kr, dr, _ = eR.extract(right, cap=nf + 200)
    return left, right, eL, eR, kl, dl, kr, dr


def _numpy_rowband(eL, eR, kl, dl, kr, dr, mbf, mb):
    """The same algorithm written the slow, obvious way."""
    f32 = np.float32
    sf, inv = eL.params()["sf"], eL.params()["inv"]
    levelsL = [eL.level(l).astype(np.int32) for l in range(8)]
    levelsR = [eR.level(l).astype(np.int32) for l in range(8)]
    nrows = levelsL[0].shape[0]
    rows = [[] for _ in range(nrows)]
    for i in range(len(kr)):
        r = f32(2.0) * sf[kr["octave"][i]]
        lo, hi = int(np.floor(f32(kr["y"][i] - r))), int(np.ceil(f32(kr["y"][i] + r)))
        for y in range(max(lo, 0), min(hi, nrows - 1) + 1):
            rows[y].append(i)
    ur = np.full(len(kl), -1, f32); dp = np.full(len(kl), -1, f32)
    maxD = f32(f32(mbf) / f32(mb))
    found = []
    for i in range(len(kl)):
        uL, vL, lv = kl["x"][i], kl["y"][i], int(kl["octave"][i])
        best, bi = 100, 0
        for j in rows[int(vL)]:
            if abs(int(kr["octave"][j]) - lv) > 1:
                continue
            if not (f32(uL - maxD) <= kr["x"][j] <= uL):
                continue
            d = int(np.unpackbits(dl[i] ^ dr[j]).sum())
            if d < best:
                best, bi = d, j
        if best >= 75:
            continue
        # std::round: half away from zero (coordinates are positive)
        cu, cv, cr = (int(np.floor(float(f32(v * inv[lv])) + 0.5)) for v in (uL, vL, kr["x"][bi]))
        IL, IR = levelsL[lv], levelsR[lv]
        if cr < 0 or cr + 11 >= IR.shape[1]:
            continue
        win = IL[cv - 5:cv + 6, cu - 5:cu + 6]
        sads = [int(np.abs(win - IR[cv - 5:cv + 6, cr + s - 5:cr + s + 6]).sum()) for s in range(-5, 6)]
        k = int(np.argmin(sads))
        if k in (0, 10):
            continue
        d1, d2, d3 = f32(sads[k - 1]), f32(sads[k]), f32(sads[k + 1])
        with np.errstate(all="ignore"):
            delta = f32(f32(d1 - d3) / f32(f32(2.0) * f32(f32(d1 + d3) - f32(f32(2.0) * d2))))
        if delta < -1 or delta > 1:
            continue
        bu = f32(sf[lv] * f32(f32(f32(cr) + f32(k - 5)) + delta))
        disp = f32(uL - bu)
        if disp >= 0 and disp < maxD:
            if disp <= 0:
                disp = f32(0.01); bu = f32(uL - f32(0.01))
            dp[i] = f32(f32(mbf) / disp); ur[i] = bu
            found.append((sads[k], i))
    if not found:
        return 0, ur, dp
    found.sort()
    th = f32(f32(f32(1.5) * f32(1.4)) * f32(found[len(found) // 2][0]))
    kept = len(found)
    for s, i in found:
        if not (f32(s) < th):
            ur[i] = -1; dp[i] = -1; kept -= 1
    return kept, ur, dp


def test_oracle_rowband_vs_numpy_restatement(oracle_mod):
    _, _, eL, eR, kl, dl, kr, dr = _pair(5, 500, 200, 600, oracle_mod)
    n, ur, dp = oracle_mod.stereo_rowband(eL, eR, kl, dl, kr, dr, FX * B, B)
    rn, rur, rdp = _numpy_rowband(eL, eR, kl, dl, kr, dr, FX * B, B)
    assert n == rn and np.array_equal(ur, rur) and np.array_equal(dp, rdp)
    assert n > 50 and (ur >= 0).sum() == n


@pytest.mark.gpu
@pytest.mark.parametrize("seed,W,H,nf", [(11, 1241, 376, 2000), (12, 752, 480, 1200), (13, 640, 480, 1000)])
def test_cuda_rowband_stereo_vs_oracle(orbx_mod, oracle_mod, seed, W, H, nf):
    """BASELINE config 2 with the classical association instead of the 2000×2000 brute force: left/right extraction, row-band Hamming
    search, SAD refinement, median cut — every mvuRight / mvDepth equals the oracle's."""
    left, right, eL, eR, kl, dl, kr, dr = _pair(seed, W, H, nf, oracle_mod)
    xL = orbx_mod.ORBextractor(nf, 1.2, 8, 20, 7, max_width=W, max_height=H)
    xR = orbx_mod.ORBextractor(nf, 1.2, 8, 20, 7, max_width=W, max_height=H)
    _, gkl, gdl = xL(left)
    _, gkr, gdr = xR(right)
    assert gkl.tobytes() == kl.tobytes() and gkr.tobytes() == kr.tobytes() and np.array_equal(gdl, dl) and np.array_equal(gdr, dr)
    n, ur, dp = orbx_mod.ComputeStereoMatches(xL, xR, gkl, gdl, gkr, gdr, FX * B, B)
    rn, rur, rdp = oracle_mod.stereo_rowband(eL, eR, kl, dl, kr, dr, FX * B, B)
    assert n == rn and np.array_equal(ur, rur) and np.array_equal(dp, rdp)
    assert n > 100
