"""cv2-backed restatement of the reference ORB extractor.  TEST INFRASTRUCTURE, NOT PRODUCT.

Calls the REAL OpenCV primitives (the container's cv2 wheel) for resize / copyMakeBorder / FAST /
GaussianBlur / fastAtan2 / BFMatcher and restates only the orchestration of
`/root/reference/src/ORBextractor.cc` (cell grid :781-869, DANI filter :871-907, quadtree :480-779,
IC_Angle :76-103, rBRIEF :107-146, output ordering :1142-1206, pyramid :1209-1234) in Python with
numpy.float32 arithmetic.  It exists to pin `liborb_oracle.so` (the C++ oracle) to real OpenCV
outputs: tests compare the two stage by stage, and `tools/gen_golden.py` uses this module to write the
fixtures in tests/golden/.

libm: cos/sin go through glibc `cosf`/`sinf` via ctypes (numpy's float32 cos/sin are a different
implementation).  std::sort tie order comes from `orc_sort_nodes` (real libstdc++ std::sort).
"""
from __future__ import annotations

import ctypes
import math
import os

import numpy as np

f32 = np.float32
EDGE = 19
HALF_PATCH = 15
PATCH = 31

_libm = ctypes.CDLL("libm.so.6")
_libm.cosf.restype = ctypes.c_float
_libm.cosf.argtypes = [ctypes.c_float]
_libm.sinf.restype = ctypes.c_float
_libm.sinf.argtypes = [ctypes.c_float]

_here = os.path.dirname(os.path.abspath(__file__))


def _pattern() -> np.ndarray:
    txt = open(os.path.join(_here, "orb_pattern.inc")).read()
    vals = [int(v) for v in txt.split("\n", 1)[1].replace("\n", "").split(",") if v.strip()]
    return np.array(vals, dtype=np.int32).reshape(256, 4)


PATTERN = _pattern()


def cv_round(v) -> int:
    """cvRound(float): round half to even."""
    return int(np.rint(f32(v)))


class Params:
    """ORBextractor ctor maths (:409-469)."""

    def __init__(self, nfeatures, scale_factor, nlevels, ini_th, min_th):
        self.nfeatures, self.nlevels, self.ini_th, self.min_th = nfeatures, nlevels, int(ini_th), int(min_th)
        sfd = float(f32(scale_factor))  # double member initialised from a float
        self.sf = [f32(1.0)]
        self.sig2 = [f32(1.0)]
        for i in range(1, nlevels):
            self.sf.append(f32(float(self.sf[i - 1]) * sfd))
            self.sig2.append(f32(self.sf[i] * self.sf[i]))
        self.inv = [f32(1.0) / s for s in self.sf]
        self.invsig2 = [f32(1.0) / s for s in self.sig2]
        factor = f32(1.0 / sfd)
        want = f32(f32(f32(nfeatures) * f32(f32(1) - factor)) / f32(f32(1) - f32(math.pow(float(factor), float(nlevels)))))
        self.quota = []
        tot = 0
        for _ in range(nlevels - 1):
            q = cv_round(want)
            self.quota.append(q)
            tot += q
            want = f32(want * factor)
        self.quota.append(max(nfeatures - tot, 0))
        umax = [0] * (HALF_PATCH + 1)
        vmax = int(math.floor(float(f32(f32(HALF_PATCH) * f32(math.sqrt(2.0)) / f32(2)) + f32(1))))
        vmin = int(math.ceil(float(f32(f32(HALF_PATCH) * f32(math.sqrt(2.0)) / f32(2)))))
        for v in range(vmax + 1):
            umax[v] = int(np.rint(math.sqrt(HALF_PATCH * HALF_PATCH - v * v)))
        v0 = 0
        for v in range(HALF_PATCH, vmin - 1, -1):
            while umax[v0] == umax[v0 + 1]:
                v0 += 1
            umax[v] = v0
            v0 += 1
        self.umax = umax


def pyramid(img: np.ndarray, P: Params):
    """ComputePyramid (:1209-1234): returns list of padded planes; ROI = plane[19:-19, 19:-19]."""
    import cv2
    planes = []
    rows, cols = img.shape
    for l in range(P.nlevels):
        s = P.inv[l]
        w, h = cv_round(f32(cols) * s), cv_round(f32(rows) * s)
        if l == 0:
            padded = cv2.copyMakeBorder(img, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101)
        else:
            prev = planes[l - 1][EDGE:-EDGE, EDGE:-EDGE]
            lvl = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR)
            padded = cv2.copyMakeBorder(lvl, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101 | cv2.BORDER_ISOLATED)
        assert padded.shape == (h + 2 * EDGE, w + 2 * EDGE)
        planes.append(padded)
    return planes


def _sort_nodes(lib, sizes, ulx):
    n = len(sizes)
    a = np.asarray(sizes, dtype=np.int32)
    b = np.asarray(ulx, dtype=np.int32)
    perm = np.empty(n, dtype=np.int32)
    lib.orc_sort_nodes(a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), n,
                       perm.ctypes.data_as(ctypes.c_void_p))
    return perm.tolist()


class _Node:
    __slots__ = ("x0", "x1", "y0", "y1", "pts", "leaf", "alive")

    def __init__(self, x0, x1, y0, y1):
        self.x0, self.x1, self.y0, self.y1 = x0, x1, y0, y1
        self.pts = []
        self.leaf = False
        self.alive = True


def _split(n: _Node, X, Y):
    hx = int(math.ceil(float(f32(n.x1 - n.x0) / f32(2))))
    hy = int(math.ceil(float(f32(n.y1 - n.y0) / f32(2))))
    xm, ym = n.x0 + hx, n.y0 + hy
    ch = [_Node(n.x0, xm, n.y0, ym), _Node(xm, n.x1, n.y0, ym), _Node(n.x0, xm, ym, n.y1), _Node(xm, n.x1, ym, n.y1)]
    fxm, fym = f32(xm), f32(ym)
    for i in n.pts:
        q = (0 if X[i] < fxm else 1) + (0 if Y[i] < fym else 2)
        ch[q].pts.append(i)
    for c in ch:
        c.leaf = len(c.pts) == 1
    return ch


def distribute(lib, X, Y, R, minX, maxX, minY, maxY, N):
    """DistributeOctTree (:555-779).  X,Y float32 arrays, R responses.  Returns selected indices in
    list order.  The std::list is a Python list with index 0 = front."""
    v = float(f32(maxX - minX) / f32(maxY - minY))  # std::round(float): half away from zero
    nIni = int(math.floor(v + 0.5)) if v >= 0 else -int(math.floor(-v + 0.5))
    hX = f32(maxX - minX) / f32(nIni)
    nodes = []
    for i in range(nIni):
        nodes.append(_Node(int(hX * f32(i)), int(hX * f32(i + 1)), 0, maxY - minY))
    roots = list(nodes)
    for i in range(len(X)):
        roots[int(X[i] / hX)].pts.append(i)
    kept = []
    for n in nodes:
        if len(n.pts) == 1:
            n.leaf = True
            kept.append(n)
        elif len(n.pts) > 1:
            kept.append(n)
    nodes = kept
    done = False
    while not done:
        before = len(nodes)
        pending = []
        n_expand = 0
        front = []  # children pushed to the front during this pass (front[0] is pushed first)
        rest = []
        for n in nodes:
            if n.leaf:
                rest.append(n)
                continue
            for c in _split(n, X, Y):
                if c.pts:
                    front.append(c)
                    if len(c.pts) > 1:
                        n_expand += 1
                        pending.append(c)
        nodes = front[::-1] + rest
        if len(nodes) >= N or len(nodes) == before:
            done = True
        elif len(nodes) + 3 * n_expand > N:
            while not done:
                before = len(nodes)
                work = pending
                pending = []
                perm = _sort_nodes(lib, [len(n.pts) for n in work], [n.x0 for n in work])
                work = [work[p] for p in perm]
                front = []
                size = len(nodes)
                for j in range(len(work) - 1, -1, -1):
                    n = work[j]
                    for c in _split(n, X, Y):
                        if c.pts:
                            front.append(c)
                            size += 1
                            if len(c.pts) > 1:
                                pending.append(c)
                    n.alive = False
                    size -= 1
                    if size >= N:
                        break
                nodes = front[::-1] + [n for n in nodes if n.alive]
                assert len(nodes) == size
                if len(nodes) >= N or len(nodes) == before:
                    done = True
    out = []
    for n in nodes:
        best = n.pts[0]
        for k in n.pts[1:]:
            if R[k] > R[best]:
                best = k
        out.append(best)
    return out


def detect_level(lib, plane: np.ndarray, level: int, P: Params, rects):
    """Cell loop + DANI filter (:781-907) on one padded plane → candidate arrays (x,y,response)."""
    import cv2
    roi = plane[EDGE:-EDGE, EDGE:-EDGE]
    h, w = roi.shape
    minBX = minBY = EDGE - 3
    maxBX, maxBY = w - EDGE + 3, h - EDGE + 3
    width, height = f32(maxBX - minBX), f32(maxBY - minBY)
    nCols, nRows = int(width / f32(35)), int(height / f32(35))
    wCell = int(math.ceil(float(width / f32(nCols))))
    hCell = int(math.ceil(float(height / f32(nRows))))
    det_ini = cv2.FastFeatureDetector_create(P.ini_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    det_min = cv2.FastFeatureDetector_create(P.min_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    X = np.empty(0, dtype=f32)
    Y = np.empty(0, dtype=f32)
    R = np.empty(0, dtype=f32)
    scale = P.sf[level]
    inv = f32(1) / scale
    for i in range(nRows):
        iniY = minBY + i * hCell
        maxY = iniY + hCell + 6
        if iniY >= maxBY - 3:
            continue
        maxY = min(maxY, maxBY)
        for j in range(nCols):
            iniX = minBX + j * wCell
            maxX = iniX + wCell + 6
            if iniX >= maxBX - 6:
                continue
            maxX = min(maxX, maxBX)
            cell = roi[iniY:maxY, iniX:maxX]
            kps = det_ini.detect(cell)
            if len(kps) == 0:
                kps = det_min.detect(cell)
            if len(kps):
                X = np.concatenate([X, np.array([k.pt[0] for k in kps], dtype=f32) + f32(j * wCell)])
                Y = np.concatenate([Y, np.array([k.pt[1] for k in kps], dtype=f32) + f32(i * hCell)])
                R = np.concatenate([R, np.array([k.response for k in kps], dtype=f32)])
            X = (X + f32(minBX)) * scale
            Y = (Y + f32(minBY)) * scale
            if len(rects) and len(X):
                px = np.rint(X).astype(np.int64)
                py = np.rint(Y).astype(np.int64)
                hit = np.zeros(len(X), dtype=bool)
                for (rx, ry, rw, rh) in rects:
                    hit |= (rx <= px) & (px < rx + rw) & (ry <= py) & (py < ry + rh)
                X, Y, R = X[~hit], Y[~hit], R[~hit]
            X = X * inv - f32(minBX)
            Y = Y * inv - f32(minBY)
    return X.astype(f32), Y.astype(f32), R.astype(f32), (minBX, maxBX, minBY, maxBY)


def ic_angle(plane: np.ndarray, x, y, umax):
    import cv2
    cx, cy = cv_round(x) + EDGE, cv_round(y) + EDGE
    m01 = m10 = 0
    row = plane[cy].astype(np.int64)
    us = np.arange(-HALF_PATCH, HALF_PATCH + 1)
    m10 += int((us * row[cx - HALF_PATCH: cx + HALF_PATCH + 1]).sum())
    for v in range(1, HALF_PATCH + 1):
        d = umax[v]
        u = np.arange(-d, d + 1)
        lo = plane[cy + v, cx - d: cx + d + 1].astype(np.int64)
        up = plane[cy - v, cx - d: cx + d + 1].astype(np.int64)
        m01 += v * int((lo - up).sum())
        m10 += int((u * (lo + up)).sum())
    return f32(cv2.fastAtan2(float(f32(m01)), float(f32(m10))))


def rbrief(blur: np.ndarray, x, y, angle_deg):
    factor_pi = f32(math.pi / float(f32(180.0)))
    ang = f32(f32(angle_deg) * factor_pi)
    a, b = f32(_libm.cosf(float(ang))), f32(_libm.sinf(float(ang)))
    cx, cy = cv_round(x), cv_round(y)
    px0, py0, px1, py1 = (PATTERN[:, k].astype(f32) for k in range(4))
    r0 = np.rint(px0 * b + py0 * a).astype(np.int64)
    c0 = np.rint(px0 * a - py0 * b).astype(np.int64)
    r1 = np.rint(px1 * b + py1 * a).astype(np.int64)
    c1 = np.rint(px1 * a - py1 * b).astype(np.int64)
    bits = (blur[cy + r0, cx + c0] < blur[cy + r1, cx + c1]).astype(np.uint8)
    return np.packbits(bits.reshape(32, 8), axis=1, bitorder="little").reshape(32)


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])


def extract(lib, img: np.ndarray, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7,
            rects=(), lap=(0, 0), taps=None):
    """ORBextractor::operator() (:1125-1207).  Returns (kps[KP_DTYPE], desc[n,32], mono_index)."""
    import cv2
    P = Params(nfeatures, scale_factor, nlevels, ini_th, min_th)
    planes = pyramid(img, P)
    per_level = []
    for l in range(nlevels):
        X, Y, R, (minBX, maxBX, minBY, maxBY) = detect_level(lib, planes[l], l, P, rects)
        sel = distribute(lib, X, Y, R, minBX, maxBX, minBY, maxBY, P.quota[l]) if len(X) else []
        k = np.zeros(len(sel), dtype=KP_DTYPE)
        k["x"] = X[sel] + f32(minBX)
        k["y"] = Y[sel] + f32(minBY)
        k["response"] = R[sel]
        k["octave"] = l
        k["size"] = f32(int(f32(PATCH) * P.sf[l]))
        k["class_id"] = -1
        for i in range(len(k)):
            k["angle"][i] = ic_angle(planes[l], k["x"][i], k["y"][i], P.umax)
        per_level.append(k)
        if taps is not None:
            taps.setdefault("cand", []).append((X, Y, R))
            taps.setdefault("sel", []).append(k.copy())
    n = sum(len(k) for k in per_level)
    out = np.zeros(n, dtype=KP_DTYPE)
    desc = np.zeros((n, 32), dtype=np.uint8)
    mono, stereo = 0, n - 1
    for l in range(nlevels):
        k = per_level[l]
        if len(k) == 0:
            if taps is not None:
                taps.setdefault("blur", []).append(None)
            continue
        work = planes[l][EDGE:-EDGE, EDGE:-EDGE].copy()
        work = cv2.GaussianBlur(work, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        if taps is not None:
            taps.setdefault("blur", []).append(work)
        scale = P.sf[l]
        for i in range(len(k)):
            d = rbrief(work, k["x"][i], k["y"][i], k["angle"][i])
            kp = k[i].copy()
            if l != 0:
                kp["x"] = f32(kp["x"] * scale)
                kp["y"] = f32(kp["y"] * scale)
            if lap[0] <= kp["x"] <= lap[1]:
                at = stereo
                stereo -= 1
            else:
                at = mono
                mono += 1
            out[at] = kp
            desc[at] = d
    if taps is not None:
        taps["planes"] = planes
        taps["params"] = P
    return out, desc, mono
