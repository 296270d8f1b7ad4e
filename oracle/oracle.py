"""ctypes binding of liborb_oracle.so (the C++ CPU oracle).  TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

_lib = None


def build(force: bool = False) -> str:
    so = os.path.join(_here, "liborb_oracle.so")
    srcs = [os.path.join(_here, f) for f in ("orb_oracle.cpp", "bow_oracle.cpp", "orb_oracle.h")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _here, "liborb_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
        L.orc_create.restype = vp
        L.orc_create.argtypes = [i32, f32, i32, i32, i32]
        L.orc_destroy.argtypes = [vp]
        L.orc_params.argtypes = [vp] * 7
        L.orc_extract.restype = i32
        L.orc_extract.argtypes = [vp, vp, i32, i32, sz, vp, i32, i32, i32, vp, vp, i32, vp, vp]
        L.orc_level_size.argtypes = [vp, i32, vp, vp]
        L.orc_get_level.argtypes = [vp, i32, i32, vp, sz]
        L.orc_get_blurred.argtypes = [vp, i32, vp, sz]
        L.orc_get_candidates.restype = i32
        L.orc_get_candidates.argtypes = [vp, i32, vp, i32]
        L.orc_get_selected.restype = i32
        L.orc_get_selected.argtypes = [vp, i32, vp, i32]
        L.orc_resize_linear_u8.argtypes = [vp, i32, i32, sz, vp, i32, i32, sz]
        L.orc_border_reflect101_u8.argtypes = [vp, i32, i32, sz, vp, sz, i32]
        L.orc_gaussian7_u8.argtypes = [vp, i32, i32, sz, vp, sz]
        L.orc_fast9_nms.restype = i32
        L.orc_fast9_nms.argtypes = [vp, i32, i32, sz, i32, vp, i32]
        L.orc_fast_atan2.restype = f32
        L.orc_fast_atan2.argtypes = [f32, f32]
        L.orc_cvround.restype = i32
        L.orc_cvround.argtypes = [f32]
        L.orc_sincos.argtypes = [f32, vp, vp]
        L.orc_sincos_array.argtypes = [vp, C.c_int64, vp, vp, i32]
        L.orc_distribute.restype = i32
        L.orc_distribute.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, i32]
        L.orc_sort_nodes.argtypes = [vp, vp, i32, vp]
        L.orc_descriptor_distance.restype = i32
        L.orc_descriptor_distance.argtypes = [vp, vp]
        L.orc_knn2.argtypes = [vp, i32, vp, C.c_int64, vp, vp, i32]
        L.orc_ratio_test.argtypes = [vp, i32, C.c_double, vp]
        L.orc_top2_lists.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp]
        L.orc_search_by_projection.restype = i32
        L.orc_search_by_projection.argtypes = [vp, vp, vp, i32, vp, vp, f32, f32, f32, f32, vp, vp, vp, vp, vp, vp, i32, f32, f32, i32, f32, vp]
        L.orc_search_by_bow.restype = i32
        L.orc_search_by_bow.argtypes = [vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, i32, vp, vp, vp, i32, f32, i32, vp]
        L.orc_search_by_bow_kf.restype = i32
        L.orc_search_by_bow_kf.argtypes = [vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, i32, vp, vp, vp, vp, i32, f32, i32, vp]
        L.orc_three_maxima.argtypes = [vp, i32, vp]
        L.orc_rot_hist_filter.argtypes = [vp, vp, i32, vp]
        L.orc_features_in_area.restype = i32
        L.orc_features_in_area.argtypes = [vp, vp, i32, f32, f32, f32, f32, vp, i32, i32, i32, vp, vp, i32]
        L.orc_stereo_rowband.restype = i32
        L.orc_stereo_rowband.argtypes = [vp, vp, vp, vp, i32, vp, vp, i32, f32, f32, vp, vp]
        L.orc_stereo_tail.restype = i32
        L.orc_stereo_tail.argtypes = [vp, vp, i32, i32, vp, vp, vp, f32, f32, vp, vp]
        L.orc_search_init.restype = i32
        L.orc_search_init.argtypes = [vp, vp, vp, i32, vp, vp, i32, vp, vp, f32, i32, vp]
        L.orc_vocab_from_nodes.restype = vp
        L.orc_vocab_from_nodes.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32]
        L.orc_vocab_load_text.restype = vp
        L.orc_vocab_load_text.argtypes = [C.c_char_p]
        L.orc_vocab_free.argtypes = [vp]
        L.orc_vocab_words.restype = i32
        L.orc_vocab_words.argtypes = [vp]
        L.orc_vocab_nodes.restype = i32
        L.orc_vocab_nodes.argtypes = [vp]
        L.orc_bow_transform.restype = i32
        L.orc_bow_transform.argtypes = [vp, vp, i32, i32] + [vp] * 9
        L.orc_undistort_points.argtypes = [vp, i32, f32, f32, f32, f32, vp, i32, vp]
        L.orc_image_bounds.argtypes = [i32, i32, f32, f32, f32, f32, vp, i32, vp]
        L.orc_bow_score_l1.restype = C.c_double
        L.orc_bow_score_l1.argtypes = [vp, vp, i32, vp, vp, i32]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Extractor:
    """Mirror of ORB_SLAM3::ORBextractor over the C++ oracle."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self.h = self.L.orc_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        if not self.h:
            raise ValueError("bad extractor parameters")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def params(self):
        n = self.nlevels
        sf, inv, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        quota = np.zeros(n, np.int32)
        umax = np.zeros(16, np.int32)
        self.L.orc_params(self.h, _p(sf), _p(inv), _p(s2), _p(is2), _p(quota), _p(umax))
        return dict(sf=sf, inv=inv, sigma2=s2, inv_sigma2=is2, quota=quota, umax=umax)

    def extract(self, img, rects=(), lap=(0, 0), cap=None):
        """→ (rc, kps, desc, mono_index) like ORBextractor::operator()."""
        img = np.asarray(img)
        assert img.dtype == np.uint8 and img.ndim == 2 and (img.size == 0 or img.strides[1] == 1)
        cap = cap or (self.nfeatures + 64 * self.nlevels + 64)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        mono = C.c_int(0)
        r = np.ascontiguousarray(np.asarray(rects, np.int32).reshape(-1, 4))
        rc = self.L.orc_extract(self.h, _p(img), img.shape[0], img.shape[1], img.strides[0] if img.size else 0,
                                _p(r), len(r), lap[0], lap[1], _p(kps), _p(desc), cap, C.byref(n), C.byref(mono))
        if rc != 0:
            return rc, None, None, mono.value
        return 0, kps[: n.value].copy(), desc[: n.value].copy(), mono.value

    def level_size(self, l):
        w, h = C.c_int(), C.c_int()
        self.L.orc_level_size(self.h, l, C.byref(w), C.byref(h))
        return w.value, h.value

    def level(self, l, padded=False):
        w, h = self.level_size(l)
        b = 38 if padded else 0
        out = np.zeros((h + b, w + b), np.uint8)
        self.L.orc_get_level(self.h, l, int(padded), _p(out), out.strides[0])
        return out

    def blurred(self, l):
        w, h = self.level_size(l)
        out = np.zeros((h, w), np.uint8)
        rc = self.L.orc_get_blurred(self.h, l, _p(out), out.strides[0])
        return out if rc == 0 else None

    def candidates(self, l):
        n = self.L.orc_get_candidates(self.h, l, None, 0)
        out = np.zeros(max(n, 1), KP_DTYPE)
        self.L.orc_get_candidates(self.h, l, _p(out), n)
        return out[:n]

    def selected(self, l):
        n = self.L.orc_get_selected(self.h, l, None, 0)
        out = np.zeros(max(n, 1), KP_DTYPE)
        self.L.orc_get_selected(self.h, l, _p(out), n)
        return out[:n]


def knn2(q, db, nthreads=1):
    q = np.ascontiguousarray(q, np.uint8)
    db = np.ascontiguousarray(db, np.uint8)
    idx = np.zeros((len(q), 2), np.int32)
    dist = np.zeros((len(q), 2), np.int32)
    lib().orc_knn2(_p(q), len(q), _p(db), len(db), _p(idx), _p(dist), nthreads)
    return idx, dist


def ratio_test(dist, ratio=0.7):
    dist = np.ascontiguousarray(dist, np.int32)
    keep = np.zeros(len(dist), np.uint8)
    lib().orc_ratio_test(_p(dist), len(dist), ratio, _p(keep))
    return keep.astype(bool)


def sincos_array(a, nthreads=8):
    a = np.ascontiguousarray(a, np.float32)
    s = np.empty_like(a)
    c = np.empty_like(a)
    lib().orc_sincos_array(_p(a), a.size, _p(s), _p(c), nthreads)
    return s, c


def fast_atan2(y, x):
    return float(lib().orc_fast_atan2(float(y), float(x)))


def sort_nodes(sizes, ulx):
    sizes = np.ascontiguousarray(sizes, np.int32)
    ulx = np.ascontiguousarray(ulx, np.int32)
    perm = np.zeros(len(sizes), np.int32)
    lib().orc_sort_nodes(_p(sizes), _p(ulx), len(sizes), _p(perm))
    return perm


def top2_lists(q, db, cand, off):
    q = np.ascontiguousarray(q, np.uint8); db = np.ascontiguousarray(db, np.uint8)
    cand = np.ascontiguousarray(cand, np.int32); off = np.ascontiguousarray(off, np.int32)
    bi, bd, si, sd = (np.zeros(len(q), np.int32) for _ in range(4))
    lib().orc_top2_lists(_p(q), len(q), _p(db), _p(cand), _p(off), _p(bi), _p(bd), _p(si), _p(sd))
    return bi, bd, si, sd


def search_by_projection(xy, octave, desc, bounds, scale_factors, mp_proj5, mp_level, mp_flags, mp_obs, mp_desc, nnratio=0.8, th=3.0,
                         far_points=False, th_far=50.0, u_right=None, kp_obs=None):
    """ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, …), Nleft == -1.  bounds = (minX, minY, maxX, maxY).
    Returns (nmatches, assigned[n])."""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2); octave = np.ascontiguousarray(octave, np.int32)
    desc = np.ascontiguousarray(desc, np.uint8); sf = np.ascontiguousarray(scale_factors, np.float32)
    p5 = np.ascontiguousarray(mp_proj5, np.float32).reshape(-1, 5); lv = np.ascontiguousarray(mp_level, np.int32)
    fl = np.ascontiguousarray(mp_flags, np.uint8); ob = np.ascontiguousarray(mp_obs, np.int32); md = np.ascontiguousarray(mp_desc, np.uint8)
    ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
    ko = None if kp_obs is None else np.ascontiguousarray(kp_obs, np.int32)
    out = np.zeros(len(xy), np.int32)
    n = lib().orc_search_by_projection(_p(xy), _p(octave), _p(desc), len(xy), None if ur is None else _p(ur), None if ko is None else _p(ko),
                                       *[float(b) for b in bounds], _p(sf), _p(p5), _p(lv), _p(fl), _p(ob), _p(md), len(p5), float(nnratio),
                                       float(th), int(far_points), float(th_far), _p(out))
    return n, out


def search_by_bow(kf_desc, kf_angle, kf_mp, kf_fv, f_desc, f_angle, f_fv, nnratio=0.7, check_ori=True):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&), Nleft == -1; fv = (nodes, off, idx) CSR.
    Returns (nmatches, assigned[n_f]) — assigned[i] = keyframe feature whose map point the frame feature received, or -1."""
    kd = np.ascontiguousarray(kf_desc, np.uint8); ka = np.ascontiguousarray(kf_angle, np.float32); km = np.ascontiguousarray(kf_mp, np.uint8)
    fd = np.ascontiguousarray(f_desc, np.uint8); fa = np.ascontiguousarray(f_angle, np.float32)
    kn, ko, ki = (np.ascontiguousarray(v, np.int32) for v in kf_fv)
    fn, fo, fi = (np.ascontiguousarray(v, np.int32) for v in f_fv)
    out = np.zeros(len(fd), np.int32)
    n = lib().orc_search_by_bow(_p(kd), _p(ka), len(kd), _p(km), _p(kn), _p(ko), _p(ki), len(kn), _p(fd), _p(fa), len(fd), _p(fn), _p(fo), _p(fi),
                                len(fn), float(nnratio), int(check_ori), _p(out))
    return n, out


def search_by_bow_kf(desc1, angle1, mp1, fv1, desc2, angle2, mp2, fv2, nnratio=0.7, check_ori=True):
    """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&) → (nmatches, matches12[n1])"""
    d1 = np.ascontiguousarray(desc1, np.uint8); a1 = np.ascontiguousarray(angle1, np.float32); m1 = np.ascontiguousarray(mp1, np.uint8)
    d2 = np.ascontiguousarray(desc2, np.uint8); a2 = np.ascontiguousarray(angle2, np.float32); m2 = np.ascontiguousarray(mp2, np.uint8)
    n1, o1, i1 = (np.ascontiguousarray(v, np.int32) for v in fv1)
    n2, o2, i2 = (np.ascontiguousarray(v, np.int32) for v in fv2)
    out = np.zeros(len(d1), np.int32)
    n = lib().orc_search_by_bow_kf(_p(d1), _p(a1), len(d1), _p(m1), _p(n1), _p(o1), _p(i1), len(n1), _p(d2), _p(a2), len(d2), _p(m2), _p(n2), _p(o2),
                                   _p(i2), len(n2), float(nnratio), int(check_ori), _p(out))
    return n, out


def three_maxima(counts):
    counts = np.ascontiguousarray(counts, np.int32)
    ind = np.zeros(3, np.int32)
    lib().orc_three_maxima(_p(counts), len(counts), _p(ind))
    return tuple(int(v) for v in ind)


def rot_hist_filter(a, b):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    keep = np.zeros(len(a), np.uint8)
    lib().orc_rot_hist_filter(_p(a), _p(b), len(a), _p(keep))
    return keep.astype(bool)


def search_init(d1, a1, o1, d2, a2, cand, off, nnratio=0.9, check_ori=True):
    d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
    a1 = np.ascontiguousarray(a1, np.float32); a2 = np.ascontiguousarray(a2, np.float32)
    o1 = np.ascontiguousarray(o1, np.int32)
    cand = np.ascontiguousarray(cand, np.int32); off = np.ascontiguousarray(off, np.int32)
    m12 = np.zeros(len(d1), np.int32)
    n = lib().orc_search_init(_p(d1), _p(a1), _p(o1), len(d1), _p(d2), _p(a2), len(d2), _p(cand), _p(off), float(nnratio), int(check_ori), _p(m12))
    return n, m12


def features_in_area(xy, octave, bounds, queries, min_level=-1, max_level=-1):
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2); octave = np.ascontiguousarray(octave, np.int32)
    queries = np.ascontiguousarray(queries, np.float32).reshape(-1, 3)
    off = np.zeros(len(queries) + 1, np.int32)
    total = lib().orc_features_in_area(_p(xy), _p(octave), len(xy), *[float(b) for b in bounds], _p(queries), len(queries), min_level, max_level, _p(off), None, 0)
    cand = np.zeros(max(total, 1), np.int32)
    lib().orc_features_in_area(_p(xy), _p(octave), len(xy), *[float(b) for b in bounds], _p(queries), len(queries), min_level, max_level, _p(off), _p(cand), total)
    return off, cand[:total]


def stereo_rowband(exL, exR, kL, dL, kR, dR, mbf, mb):
    """Classical ComputeStereoMatches over the pyramids the two oracle extractors hold from their last extract() → (n, mvuRight, mvDepth)."""
    kL = np.ascontiguousarray(kL, KP_DTYPE); kR = np.ascontiguousarray(kR, KP_DTYPE)
    dL = np.ascontiguousarray(dL, np.uint8); dR = np.ascontiguousarray(dR, np.uint8)
    ur = np.zeros(len(kL), np.float32); dp = np.zeros(len(kL), np.float32)
    n = lib().orc_stereo_rowband(exL.h, exR.h, _p(kL), _p(dL), len(kL), _p(kR), _p(dR), len(kR), float(mbf), float(mb), _p(ur), _p(dp))
    return n, ur, dp


def stereo_tail(uL, uR, idx, dist, keep, mbf, mb):
    uL = np.ascontiguousarray(uL, np.float32); uR = np.ascontiguousarray(uR, np.float32)
    idx = np.ascontiguousarray(idx, np.int32); dist = np.ascontiguousarray(dist, np.int32); keep = np.ascontiguousarray(keep, np.uint8)
    ur = np.zeros(len(uL), np.float32); dp = np.zeros(len(uL), np.float32)
    n = lib().orc_stereo_tail(_p(uL), _p(uR), len(uL), len(uR), _p(idx), _p(dist), _p(keep), float(mbf), float(mb), _p(ur), _p(dp))
    return n, ur, dp


class Vocabulary:
    """DBoW2 vocabulary tree + transform (bow_oracle.cpp).  Build from a node stream (dict from synth.vocabulary) or a
    text file in the reference's ORBvoc.txt format."""

    def __init__(self, voc=None, path=None):
        L = lib()
        if path is not None:
            self.h = L.orc_vocab_load_text(str(path).encode())
        else:
            par = np.ascontiguousarray(voc["parent"], np.int32); leaf = np.ascontiguousarray(voc["is_leaf"], np.uint8)
            desc = np.ascontiguousarray(voc["desc"], np.uint8); w = np.ascontiguousarray(voc["weight"], np.float64)
            self.h = L.orc_vocab_from_nodes(_p(par), _p(leaf), _p(desc), _p(w), len(par), int(voc["k"]), int(voc["L"]),
                                            int(voc["scoring"]), int(voc["weighting"]))
        if not self.h:
            raise ValueError("bad vocabulary")

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_vocab_free(self.h)
            self.h = None

    @property
    def n_words(self):
        return lib().orc_vocab_words(self.h)

    def transform(self, desc, levelsup=4):
        """→ dict(word_id[n], node_id[n], bow_ids, bow_vals, fv_nodes, fv_off, fv_idx) in std::map order."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        m = max(n, 1)
        wid = np.zeros(m, np.uint32); nid = np.zeros(m, np.uint32)
        bi = np.zeros(m, np.uint32); bv = np.zeros(m, np.float64)
        fn = np.zeros(m, np.uint32); fo = np.zeros(m + 1, np.int32); fi = np.zeros(m, np.uint32)
        nb, nf = C.c_int(0), C.c_int(0)
        lib().orc_bow_transform(self.h, _p(desc), n, int(levelsup), _p(wid), _p(nid), _p(bi), _p(bv), C.byref(nb), _p(fn), _p(fo), _p(fi), C.byref(nf))
        nb, nf = nb.value, nf.value
        return dict(word_id=wid[:n], node_id=nid[:n], bow_ids=bi[:nb], bow_vals=bv[:nb], fv_nodes=fn[:nf], fv_off=fo[:nf + 1], fv_idx=fi[:fo[nf]])


def bow_score_l1(ids1, v1, ids2, v2):
    ids1 = np.ascontiguousarray(ids1, np.uint32); ids2 = np.ascontiguousarray(ids2, np.uint32)
    v1 = np.ascontiguousarray(v1, np.float64); v2 = np.ascontiguousarray(v2, np.float64)
    return lib().orc_bow_score_l1(_p(ids1), _p(v1), len(ids1), _p(ids2), _p(v2), len(ids2))


def undistort_points(xy, fx, fy, cx, cy, D):
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2); D = np.ascontiguousarray(D, np.float32)
    out = np.zeros_like(xy)
    lib().orc_undistort_points(_p(xy), len(xy), float(fx), float(fy), float(cx), float(cy), _p(D), len(D), _p(out))
    return out


def image_bounds(cols, rows, fx, fy, cx, cy, D):
    D = np.ascontiguousarray(D, np.float32)
    b = np.zeros(4, np.float32)
    lib().orc_image_bounds(int(cols), int(rows), float(fx), float(fy), float(cx), float(cy), _p(D), len(D), _p(b))
    return b
