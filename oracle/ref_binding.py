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
