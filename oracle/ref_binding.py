"""ctypes binding of oracle/_ref/libref_orb.so — the reference's OWN src/ORBextractor.cc compiled unmodified
against oracle/shim (see oracle/Makefile target `ref`).  TEST INFRASTRUCTURE / CPU BASELINE, NOT PRODUCT.

The .so only exists where /root/reference was available at build time (the build container); it travels to
the GPU box with the repo snapshot.  `available()` says whether it can be used.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .oracle import KP_DTYPE, _p, build as _build_oracle

_here = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_here, "_ref", "libref_orb.so")
_lib = None


def available() -> bool:
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        _build_oracle()  # libref_orb.so links liborb_oracle.so (rpath $ORIGIN/..)
        L = C.CDLL(SO)
        vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
        L.ref_create.restype = vp
        L.ref_create.argtypes = [i32, f32, i32, i32, i32]
        L.ref_destroy.argtypes = [vp]
        L.ref_extract.restype = i32
        L.ref_extract.argtypes = [vp, vp, i32, i32, sz, vp, i32, i32, i32, vp, vp, i32, vp, vp]
        L.ref_level.restype = i32
        L.ref_level.argtypes = [vp, i32, vp, sz, vp, vp]
        _lib = L
    return _lib


class Extractor:
    """The reference's ORB_SLAM3::ORBextractor itself (same call shape as oracle.Extractor)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.h = self.L.ref_create(nfeatures, scale_factor, nlevels, ini_th, min_th)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_destroy(self.h)
            self.h = None

    def extract(self, img, rects=(), lap=(0, 0), cap=None):
        img = np.asarray(img)
        cap = cap or (self.nfeatures + 64 * self.nlevels + 64)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n, mono = C.c_int(0), C.c_int(0)
        r = np.ascontiguousarray(np.asarray(rects, np.int32).reshape(-1, 4))
        rc = self.L.ref_extract(self.h, _p(img), img.shape[0], img.shape[1], img.strides[0] if img.size else 0, _p(r), len(r),
                                lap[0], lap[1], _p(kps), _p(desc), cap, C.byref(n), C.byref(mono))
        if rc != 0:
            return rc, None, None, mono.value
        return 0, kps[: n.value].copy(), desc[: n.value].copy(), mono.value

    def level(self, l):
        w, h = C.c_int(), C.c_int()
        self.L.ref_level(self.h, l, None, 0, C.byref(w), C.byref(h))
        out = np.zeros((h.value, w.value), np.uint8)
        self.L.ref_level(self.h, l, _p(out), out.strides[0], None, None)
        return out


# ---- the reference's vendored DBoW2 (oracle/_ref/libref_bow.so) ----
SO_BOW = os.path.join(_here, "_ref", "libref_bow.so")
_lib_bow = None


def bow_available() -> bool:
    return os.path.exists(SO_BOW)


def lib_bow():
    global _lib_bow
    if _lib_bow is None:
        _build_oracle()
        L = C.CDLL(SO_BOW)
        vp, i32 = C.c_void_p, C.c_int
        L.ref_vocab_load_text.restype = vp
        L.ref_vocab_load_text.argtypes = [C.c_char_p]
        L.ref_vocab_free.argtypes = [vp]
        L.ref_vocab_size.restype = i32
        L.ref_vocab_size.argtypes = [vp]
        L.ref_bow_transform.restype = i32
        L.ref_bow_transform.argtypes = [vp, vp, i32, i32] + [vp] * 7
        L.ref_bow_words.argtypes = [vp, vp, i32, vp]
        L.ref_bow_score.restype = C.c_double
        L.ref_bow_score.argtypes = [vp, vp, vp, i32, vp, vp, i32]
        _lib_bow = L
    return _lib_bow


class Vocabulary:
    """The reference's ORBVocabulary (DBoW2::TemplatedVocabulary<FORB>) loaded with its own loadFromTextFile."""

    def __init__(self, path):
        self.L = lib_bow()
        self.h = self.L.ref_vocab_load_text(str(path).encode())
        if not self.h:
            raise ValueError("reference loadFromTextFile failed")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_vocab_free(self.h)
            self.h = None

    @property
    def n_words(self):
        return self.L.ref_vocab_size(self.h)

    def transform(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        m = max(n, 1)
        bi = np.zeros(m, np.uint32); bv = np.zeros(m, np.float64)
        fn = np.zeros(m, np.uint32); fo = np.zeros(m + 1, np.int32); fi = np.zeros(m, np.uint32)
        wid = np.zeros(m, np.uint32)
        nb, nf = C.c_int(0), C.c_int(0)
        self.L.ref_bow_transform(self.h, _p(desc), n, int(levelsup), _p(bi), _p(bv), C.byref(nb), _p(fn), _p(fo), _p(fi), C.byref(nf))
        if n:
            self.L.ref_bow_words(self.h, _p(desc), n, _p(wid))
        nb, nf = nb.value, nf.value
        return dict(word_id=wid[:n], bow_ids=bi[:nb], bow_vals=bv[:nb], fv_nodes=fn[:nf], fv_off=fo[:nf + 1], fv_idx=fi[:fo[nf]])

    def score(self, ids1, v1, ids2, v2):
        ids1 = np.ascontiguousarray(ids1, np.uint32); ids2 = np.ascontiguousarray(ids2, np.uint32)
        v1 = np.ascontiguousarray(v1, np.float64); v2 = np.ascontiguousarray(v2, np.float64)
        return self.L.ref_bow_score(self.h, _p(ids1), _p(v1), len(ids1), _p(ids2), _p(v2), len(ids2))


# ---- the reference's own matcher / Frame-grid functions, by line range (oracle/_ref/libref_match.so) ----
SO_MATCH = os.path.join(_here, "_ref", "libref_match.so")
_lib_match = None


def match_available() -> bool:
    return os.path.exists(SO_MATCH)


def lib_match():
    global _lib_match
    if _lib_match is None:
        _build_oracle()
        L = C.CDLL(SO_MATCH)
        vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
        L.refm_constants.argtypes = [vp, vp, vp]
        L.refm_descriptor_distance.restype = i32
        L.refm_descriptor_distance.argtypes = [vp, vp]
        L.refm_three_maxima.argtypes = [vp, i32, vp]
        L.refm_features_in_area.restype = i32
        L.refm_features_in_area.argtypes = [vp, i32, vp, vp, i32, i32, i32, vp, vp, i32]
        L.refm_search_init.restype = i32
        L.refm_search_init.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp, i32, f32, i32, vp]
        L.refm_search_by_projection.restype = i32
        L.refm_search_by_projection.argtypes = [vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, i32, f32, f32, i32, f32, vp]
        L.refm_search_by_bow.restype = i32
        L.refm_search_by_bow.argtypes = [vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, i32, vp, vp, vp, i32, f32, i32, vp]
        L.refm_search_by_bow_kf.restype = i32
        L.refm_search_by_bow_kf.argtypes = [vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, i32, vp, vp, vp, vp, i32, f32, i32, vp]
        L.refm_stereo_tail.restype = i32
        L.refm_stereo_tail.argtypes = [vp, i32, vp, i32, vp, vp, vp, i32, f32, f32, vp, vp]
        _lib_match = L
    return _lib_match


def _kps(kps):
    kps = np.ascontiguousarray(kps)
    assert kps.dtype == KP_DTYPE
    return kps


def _bounds(bounds):
    return np.ascontiguousarray(bounds, np.float32)          # (minX, minY, maxX, maxY)


def match_constants():
    a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
    lib_match().refm_constants(C.byref(a), C.byref(b), C.byref(c))
    return dict(TH_LOW=a.value, TH_HIGH=b.value, HISTO_LENGTH=c.value)


def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib_match().refm_descriptor_distance(_p(a), _p(b))


def three_maxima(counts):
    counts = np.ascontiguousarray(counts, np.int32)
    ind = np.zeros(3, np.int32)
    lib_match().refm_three_maxima(_p(counts), len(counts), _p(ind))
    return tuple(int(v) for v in ind)


def features_in_area(kps, bounds, queries, min_level=-1, max_level=-1):
    kps = _kps(kps); b = _bounds(bounds)
    q = np.ascontiguousarray(queries, np.float32).reshape(-1, 3)
    off = np.zeros(len(q) + 1, np.int32)
    total = lib_match().refm_features_in_area(_p(kps), len(kps), _p(b), _p(q), len(q), min_level, max_level, _p(off), None, 0)
    cand = np.zeros(max(total, 1), np.int32)
    lib_match().refm_features_in_area(_p(kps), len(kps), _p(b), _p(q), len(q), min_level, max_level, _p(off), _p(cand), total)
    return off, cand[:total]


def search_for_initialization(kps1, desc1, kps2, desc2, bounds, prev_xy, window=100, nnratio=0.9, check_ori=True):
    kps1 = _kps(kps1); kps2 = _kps(kps2); b = _bounds(bounds)
    d1 = np.ascontiguousarray(desc1, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
    prev = np.array(prev_xy, np.float32).reshape(-1, 2).copy()
    m12 = np.zeros(len(kps1), np.int32)
    n = lib_match().refm_search_init(_p(kps1), _p(d1), len(kps1), _p(kps2), _p(d2), len(kps2), _p(b), _p(prev), int(window), float(nnratio),
                                     int(check_ori), _p(m12))
    return n, m12, prev


def search_by_projection(kps, desc, bounds, scale_factors, mp_proj5, mp_level, mp_flags, mp_obs, mp_desc, nnratio=0.8, th=3.0, far_points=False,
                         th_far=50.0, u_right=None, kp_obs=None):
    kps = _kps(kps); b = _bounds(bounds); desc = np.ascontiguousarray(desc, np.uint8)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    p5 = np.ascontiguousarray(mp_proj5, np.float32).reshape(-1, 5); lv = np.ascontiguousarray(mp_level, np.int32)
    fl = np.ascontiguousarray(mp_flags, np.uint8); ob = np.ascontiguousarray(mp_obs, np.int32); md = np.ascontiguousarray(mp_desc, np.uint8)
    ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
    ko = np.full(len(kps), -1, np.int32) if kp_obs is None else np.ascontiguousarray(kp_obs, np.int32)
    out = np.zeros(len(kps), np.int32)
    n = lib_match().refm_search_by_projection(_p(kps), _p(desc), len(kps), None if ur is None else _p(ur), _p(ko), _p(b), _p(sf), len(sf), _p(p5),
                                              _p(lv), _p(fl), _p(ob), _p(md), len(p5), float(nnratio), float(th), int(far_points), float(th_far),
                                              _p(out))
    return n, out


def search_by_bow(kf_kps, kf_desc, kf_mp, kf_fv, f_kps, f_desc, f_fv, nnratio=0.7, check_ori=True):
    """The reference's own SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (src/ORBmatcher.cc:222-425); fv = (nodes, off, idx) CSR."""
    kk = _kps(kf_kps); fk = _kps(f_kps)
    kd = np.ascontiguousarray(kf_desc, np.uint8); km = np.ascontiguousarray(kf_mp, np.uint8); fd = np.ascontiguousarray(f_desc, np.uint8)
    kn, ko, ki = (np.ascontiguousarray(v, np.int32) for v in kf_fv)
    fn, fo, fi = (np.ascontiguousarray(v, np.int32) for v in f_fv)
    out = np.zeros(len(fk), np.int32)
    n = lib_match().refm_search_by_bow(_p(kk), _p(kd), len(kk), _p(km), _p(kn), _p(ko), _p(ki), len(kn), _p(fk), _p(fd), len(fk), _p(fn), _p(fo),
                                       _p(fi), len(fn), float(nnratio), int(check_ori), _p(out))
    return n, out


def search_by_bow_kf(kps1, desc1, mp1, fv1, kps2, desc2, mp2, fv2, nnratio=0.7, check_ori=True):
    """The reference's own SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&) (src/ORBmatcher.cc:760-901)"""
    k1 = _kps(kps1); k2 = _kps(kps2)
    d1 = np.ascontiguousarray(desc1, np.uint8); m1 = np.ascontiguousarray(mp1, np.uint8)
    d2 = np.ascontiguousarray(desc2, np.uint8); m2 = np.ascontiguousarray(mp2, np.uint8)
    n1, o1, i1 = (np.ascontiguousarray(v, np.int32) for v in fv1)
    n2, o2, i2 = (np.ascontiguousarray(v, np.int32) for v in fv2)
    out = np.zeros(len(k1), np.int32)
    n = lib_match().refm_search_by_bow_kf(_p(k1), _p(d1), len(k1), _p(m1), _p(n1), _p(o1), _p(i1), len(n1), _p(k2), _p(d2), len(k2), _p(m2), _p(n2),
                                          _p(o2), _p(i2), len(n2), float(nnratio), int(check_ori), _p(out))
    return n, out


def stereo_tail(u_left, u_right, iL, iR, match_distance, mbf, mb):
    uL = np.ascontiguousarray(u_left, np.float32); uR = np.ascontiguousarray(u_right, np.float32)
    iL = np.ascontiguousarray(iL, np.int32); iR = np.ascontiguousarray(iR, np.int32); md = np.ascontiguousarray(match_distance, np.float32)
    ur = np.zeros(len(uL), np.float32); dp = np.zeros(len(uL), np.float32)
    n = lib_match().refm_stereo_tail(_p(uL), len(uL), _p(uR), len(uR), _p(iL), _p(iR), _p(md), len(iL), float(mbf), float(mb), _p(ur), _p(dp))
    return n, ur, dp
