"""Developer script: stage-by-stage CUDA-vs-oracle comparison with verbose diagnostics (run under gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle
from dani_slam_b200 import orbx, synth

def compare(img, nfeat=1000, rects=(), lap=(0, 0), label=""):
    H, W = img.shape
    ex = orbx.ORBextractor(nfeat, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=4)
    ex.mvDynamicArea = list(rects)
    ref = oracle.Extractor(nfeat, 1.2, 8, 20, 7)
    rc, rk, rd, rmono = ref.extract(img, rects=rects, lap=lap, cap=nfeat + 200)
    mono, k, d = ex(img, None, lap)
    ok = True
    for l in range(8):
        a, b = ex.mvImagePyramid(l), ref.level(l)
        e1 = np.array_equal(a, b)
        ap, bp = ex.mvImagePyramid(l, padded=True), ref.level(l, True)
        e1p = np.array_equal(ap, bp)
        c, rc_ = ex.candidates(l), ref.candidates(l)
        e2 = c.tobytes() == rc_.tobytes()
        s, rs = ex.selected(l), ref.selected(l)
        rs2 = rs.copy(); rs2['angle'] = -1
        e3 = s.tobytes() == rs2.tobytes()
        bl, rbl = ex.blurred(l), ref.blurred(l)
        e4 = rbl is None or np.array_equal(bl, rbl)
        print(f"{label} L{l}: pyr {e1} pad {e1p} cand {e2} ({len(c)} vs {len(rc_)}) sel {e3} ({len(s)} vs {len(rs)}) blur {e4}")
        if not e2 and len(c) == len(rc_):
            bad = np.nonzero([c[i].tobytes() != rc_[i].tobytes() for i in range(len(c))])[0]
            print("   first cand diffs", bad[:5], c[bad[:3]], rc_[bad[:3]])
        if not e3:
            n = min(len(s), len(rs2))
            bad = np.nonzero([s[i].tobytes() != rs2[i].tobytes() for i in range(n)])[0]
            print("   first sel diffs", bad[:5], s[bad[:3]], rs2[bad[:3]])
        ok &= e1 and e1p and e2 and e3 and e4
    ek = k.tobytes() == rk.tobytes()
    ed = np.array_equal(d, rd)
    print(f"{label} final: n {len(k)} vs {len(rk)} mono {mono} vs {rmono} kps {ek} desc {ed}")
    if not ek and len(k) == len(rk):
        for f in k.dtype.names:
            if not np.array_equal(k[f], rk[f]):
                bad = np.nonzero(k[f] != rk[f])[0]
                print("   field", f, "diffs", len(bad), bad[:5], k[f][bad[:5]], rk[f][bad[:5]])
    if not ed and len(k) == len(rk):
        bad = np.nonzero((d != rd).any(axis=1))[0]
        print("   desc rows differing", len(bad), bad[:10])
    return ok and ek and ed and mono == rmono

if __name__ == "__main__":
    res = []
    res.append(compare(synth.parity_frame(3), label="parity640"))
    res.append(compare(synth.throughput_frame(0), lap=(0, 1000), label="thr640mono"))
    res.append(compare(synth.parity_frame(5, 641, 479), rects=[(100, 80, 120, 90), (300, 200, 50, 50), (10, 400, 600, 30)], label="odd+rects"))
    res.append(compare(synth.throughput_frame(1, 1241, 376), nfeat=2000, label="kitti"))
    res.append(compare(synth.throughput_frame(2, 752, 480), nfeat=1200, lap=(200, 500), label="euroc-lap"))
    print("ALL OK" if all(res) else "MISMATCH", res)
    # quick timing
    import ctypes
    imgs = np.stack([synth.throughput_frame(i) for i in range(64)])
    ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=64)
    for it in range(3):
        t = time.time(); n, mono, kps, desc = ex.extract_batch(imgs, (0, 0)); dt = time.time() - t
        print("batch64 e2e", dt, "s ->", 64 / dt, "fps; n range", n.min(), n.max())
