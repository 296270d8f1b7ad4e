"""ctypes host-side mirror of the reference's ORBextractor / ORBmatcher over liborbx.so.

The class and method names follow `/root/reference/include/ORBextractor.h:43-110` and
`include/ORBmatcher.h:36-103` so that tests read like calls into the reference.  All compute happens in
the hand-written sm_100a CUDA library behind the C ABI of `include/orbx.h`; there is no CPU fallback —
if the library is missing or no GPU is visible, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_pkg = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_pkg, "liborbx.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

OK, EMPTY, ERR_CAPACITY, ERR_GEOMETRY, ERR_ARG, ERR_CUDA = 0, -1, -2, -3, -4, -5


class OrbxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"liborbx error {code}: {msg}")
        self.code = code


_lib = None


def lib() -> C.CDLL:
    """Load liborbx.so (built by `make -C dani_slam_b200/csrc` or `__graft_entry__.build()`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OrbxError(ERR_CUDA, f"{LIB_PATH} not built: run __graft_entry__.build() (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, sz, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t, C.c_double
    L.orbx_device_count.restype = i32
    L.orbx_create.restype = vp
    L.orbx_create.argtypes = [i32, f32, i32, i32, i32, i32, i32, i32, i32]
    L.orbx_destroy.argtypes = [vp]
    L.orbx_last_error.restype = C.c_char_p
    L.orbx_last_error.argtypes = [vp]
    L.orbx_params.argtypes = [vp] * 6
    L.orbx_extract.argtypes = [vp, vp, i32, i32, sz, vp, i32, i32, i32, vp, vp, i32, vp, vp]
    L.orbx_extract_batch.argtypes = [vp, vp, i32, i32, i32, sz, vp, i32, i32, i32, vp, vp, i32, vp, vp]
    L.orbx_extract_batch_device.argtypes = [vp, vp, sz, i32, i32, i32, sz, vp, i32, i32, i32, vp, vp, i32, vp, vp]
    L.orbx_sync.argtypes = [vp]
    L.orbx_stream.restype = vp
    L.orbx_stream.argtypes = [vp]
    L.orbx_launch_count.restype = C.c_longlong
    L.orbx_launch_count.argtypes = [vp]
    L.orbx_set_profiling.argtypes = [vp, i32]
    L.orbx_get_stage_ms.argtypes = [vp, vp, vp]
    L.orbx_reset_stage_ms.argtypes = [vp]
    L.orbx_level_size.argtypes = [vp, i32, vp, vp]
    L.orbx_get_pyramid.argtypes = [vp, i32, i32, i32, vp, sz]
    L.orbx_get_blurred.argtypes = [vp, i32, i32, vp, sz]
    L.orbx_get_candidates.argtypes = [vp, i32, i32, vp, i32]
    L.orbx_get_selected.argtypes = [vp, i32, i32, vp, i32]
    L.orbx_host_alloc.restype = vp
    L.orbx_host_alloc.argtypes = [sz]
    L.orbx_host_alloc_wc.restype = vp
    L.orbx_host_alloc_wc.argtypes = [sz]
    L.orbx_host_free.argtypes = [vp]
    L.orbx_matcher_create.restype = vp
    L.orbx_matcher_create.argtypes = [i32]
    L.orbx_matcher_destroy.argtypes = [vp]
    L.orbx_matcher_last_error.restype = C.c_char_p
    L.orbx_matcher_last_error.argtypes = [vp]
    L.orbx_matcher_stream.restype = vp
    L.orbx_matcher_stream.argtypes = [vp]
    L.orbx_matcher_sync.argtypes = [vp]
    L.orbx_hamming_knn2.argtypes = [vp, vp, i32, vp, i64, vp, vp]
    L.orbx_debug_knn_tc_launches.restype = C.c_longlong
    L.orbx_debug_knn_tc_launches.argtypes = [vp]
    L.orbx_hamming_knn2_device.argtypes = [vp, vp, i32, vp, i64, i64, vp, vp]
    L.orbx_knn2_merge_device.argtypes = [vp, vp, vp, i32, i32, vp, vp]
    L.orbx_knn2_merge_packed_device.argtypes = [vp, vp, i32, i32, vp, vp]
    L.orbx_stereo_matches.argtypes = [vp, vp, vp, vp, i32, vp, vp, i32, f32, f32, vp, vp, vp]
    L.orbx_copy_only_batch.argtypes = [vp, vp, i32, i32, i32, sz, vp, vp, i32]
    L.orbx_comm_unique_id.argtypes = [vp]
    L.orbx_comm_create.restype = vp
    L.orbx_comm_create.argtypes = [i32, i32, vp, i32]
    L.orbx_comm_create_all.argtypes = [vp, i32, vp]
    L.orbx_comm_destroy.argtypes = [vp]
    L.orbx_comm_last_error.restype = C.c_char_p
    L.orbx_comm_last_error.argtypes = [vp]
    L.orbx_comm_rank.argtypes = [vp]
    L.orbx_comm_world.argtypes = [vp]
    L.orbx_knn2_sharded.argtypes = [vp, vp, vp, i32, vp, i64, i64, vp, vp]
    L.orbx_knn2_sharded_all.argtypes = [vp, vp, i32, vp, i32, vp, vp, vp, vp, vp]
    L.orbx_extract_batch_multi.argtypes = [vp, i32, vp, i32, i32, i32, sz, vp, i32, i32, i32, vp, vp, i32, vp, vp]
    L.orbx_ratio_test.argtypes = [vp, vp, i32, f64, vp]
    L.orbx_ratio_test_device.argtypes = [vp, vp, i32, f64, vp]
    L.orbx_hamming_top2_lists.argtypes = [vp, vp, i32, vp, i64, vp, vp, vp, vp, vp, vp]
    L.orbx_search_by_projection.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, i32, f32, f32, i32, f32, vp, vp]
    L.orbx_search_by_bow.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, i32, vp, vp, vp, i32, f32, i32, vp, vp]
    L.orbx_search_by_bow_keyframes.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, i32, vp, vp, vp, vp, i32, f32, i32, vp, vp]
    L.orbx_search_for_initialization_frames.argtypes = [vp, vp, vp, i32, vp, vp, i32, vp, vp, i32, f32, i32, vp, vp]
    L.orbx_rot_hist_filter.argtypes = [vp, vp, vp, i32, vp]
    L.orbx_features_in_area.argtypes = [vp, vp, vp, i32, f32, f32, f32, f32, vp, i32, i32, i32, vp, vp, i32, vp]
    L.orbx_stereo_tail.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, f32, f32, vp, vp, vp]
    L.orbx_search_for_initialization.argtypes = [vp, vp, vp, vp, i32, vp, vp, i32, vp, vp, f32, i32, vp, vp]
    L.orbx_rot_hist_filter_device.argtypes = [vp, vp, vp, i32, vp]
    L.orbx_undistort_keypoints.argtypes = [vp, vp, i32, f32, f32, f32, f32, vp, i32, vp]
    L.orbx_undistort_points_device.argtypes = [vp, vp, i32, i32, f32, f32, f32, f32, vp, i32, vp, i32]
    L.orbx_image_bounds.argtypes = [vp, i32, i32, f32, f32, f32, f32, vp, i32, vp]
    L.orbx_descriptor_distance.restype = i32
    L.orbx_descriptor_distance.argtypes = [vp, vp]
    L.orbx_vocab_create_from_nodes.restype = vp
    L.orbx_vocab_create_from_nodes.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32]
    L.orbx_vocab_load_text.restype = vp
    L.orbx_vocab_load_text.argtypes = [C.c_char_p, i32]
    L.orbx_vocab_destroy.argtypes = [vp]
    L.orbx_vocab_last_error.restype = C.c_char_p
    L.orbx_vocab_last_error.argtypes = [vp]
    L.orbx_vocab_info.argtypes = [vp, vp, vp, vp, vp]
    L.orbx_vocab_stream.restype = vp
    L.orbx_vocab_stream.argtypes = [vp]
    L.orbx_vocab_sync.argtypes = [vp]
    L.orbx_bow_transform.argtypes = [vp, vp, i32, i32] + [vp] * 9
    L.orbx_bow_transform_batch_device.argtypes = [vp, vp, C.c_size_t, vp, i32, i32, i32] + [vp] * 9
    L.orbx_debug_sort_nodes.argtypes = [vp, vp, i32, vp]
    L.orbx_debug_sort_nodes_device.argtypes = [i32, vp, vp, i32, vp]
    L.orbx_debug_sincos_device.argtypes = [i32, vp, i32, vp, vp]
    L.orbx_debug_atan2_device.argtypes = [i32, vp, vp, i32, vp]
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class ORBextractor:
    """`ORB_SLAM3::ORBextractor` (include/ORBextractor.h:43-110) on one B200.

    `ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)`; calling the object is
    `operator()(image, mask, keypoints, descriptors, vLappingArea)` and returns
    `(monoIndex, keypoints, descriptors)`.  `mvDynamicArea` is the public rect list of the reference.
    """

    HARRIS_SCORE, FAST_SCORE = 0, 1

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device=0, max_width=640,
                 max_height=480, max_batch=1):
        self.L = lib()
        self.nfeatures, self.scaleFactor, self.nlevels = int(nfeatures), float(scaleFactor), int(nlevels)
        self.iniThFAST, self.minThFAST = int(iniThFAST), int(minThFAST)  # floats are truncated like the C++ ctor
        self.device, self.max_batch = device, max_batch
        self.mvDynamicArea: list = []
        self.h = self.L.orbx_create(self.nfeatures, self.scaleFactor, self.nlevels, self.iniThFAST,
                                    self.minThFAST, device, max_width, max_height, max_batch)
        if not self.h:
            raise OrbxError(ERR_CUDA, self.L.orbx_last_error(None).decode())
        self.cap = self.nfeatures + 8 * self.nlevels + 64

    def close(self):
        if getattr(self, "h", None):
            self.L.orbx_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _err(self, rc):
        return OrbxError(rc, self.L.orbx_last_error(self.h).decode())

    # ---- getters (include/ORBextractor.h:58-78)
    def _params(self):
        n = self.nlevels
        sf, inv, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        q = np.zeros(n, np.int32)
        self.L.orbx_params(self.h, _p(sf), _p(inv), _p(s2), _p(is2), _p(q))
        return sf, inv, s2, is2, q

    def GetLevels(self):
        return self.nlevels

    def GetScaleFactor(self):
        return self.scaleFactor

    def GetScaleFactors(self):
        return self._params()[0]

    def GetInverseScaleFactors(self):
        return self._params()[1]

    def GetScaleSigmaSquares(self):
        return self._params()[2]

    def GetInverseScaleSigmaSquares(self):
        return self._params()[3]

    def features_per_level(self):
        return self._params()[4]

    def _rects(self):
        r = np.ascontiguousarray(np.asarray(self.mvDynamicArea, np.int32).reshape(-1, 4))
        return r, len(r)

    # ---- operator() (src/ORBextractor.cc:1125-1207)
    def __call__(self, image, mask=None, vLappingArea=(0, 0)):
        image = np.asarray(image)
        if image.size == 0:
            return -1, np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2, "CV_8UC1 expected (reference asserts, :1133)"
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        kps = np.zeros(self.cap, KP_DTYPE)
        desc = np.zeros((self.cap, 32), np.uint8)
        n, mono = C.c_int(0), C.c_int(0)
        r, nr = self._rects()
        rc = self.L.orbx_extract(self.h, _p(image), image.shape[0], image.shape[1], image.strides[0], _p(r), nr,
                                 int(vLappingArea[0]), int(vLappingArea[1]), _p(kps), _p(desc), self.cap,
                                 C.byref(n), C.byref(mono))
        if rc == EMPTY:
            return -1, np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        if rc != OK:
            raise self._err(rc)
        return mono.value, kps[: n.value].copy(), desc[: n.value].copy()

    def extract_batch(self, images, vLappingArea=(0, 0)):
        """Frame-batch form: images = array (B, H, W) uint8 (host).  Returns (n[B], mono[B], kps[B,cap], desc[B,cap,32])."""
        images = np.ascontiguousarray(images)
        assert images.dtype == np.uint8 and images.ndim == 3
        B, H, W = images.shape
        ptrs = (C.c_void_p * B)(*[images.ctypes.data + b * H * W for b in range(B)])
        kps = np.zeros((B, self.cap), KP_DTYPE)
        desc = np.zeros((B, self.cap, 32), np.uint8)
        n = np.zeros(B, np.int32)
        mono = np.zeros(B, np.int32)
        r, nr = self._rects()
        rc = self.L.orbx_extract_batch(self.h, ptrs, B, H, W, W, _p(r), nr, int(vLappingArea[0]),
                                       int(vLappingArea[1]), _p(kps), _p(desc), self.cap, _p(n), _p(mono))
        if rc != OK:
            raise self._err(rc)
        return n, mono, kps, desc

    def extract_batch_device(self, d_images_ptr, frame_stride, B, H, W, step, d_kps_ptr, d_desc_ptr, cap, d_n_ptr,
                             d_mono_ptr, vLappingArea=(0, 0)):
        """Device-resident batch on raw device pointers (ints); asynchronous on `stream()`."""
        r, nr = self._rects()
        rc = self.L.orbx_extract_batch_device(self.h, d_images_ptr, frame_stride, B, H, W, step, _p(r), nr,
                                              int(vLappingArea[0]), int(vLappingArea[1]), d_kps_ptr, d_desc_ptr, cap,
                                              d_n_ptr, d_mono_ptr)
        if rc != OK:
            raise self._err(rc)

    def sync(self):
        rc = self.L.orbx_sync(self.h)
        if rc != OK:
            raise self._err(rc)

    def stream(self):
        return self.L.orbx_stream(self.h)

    def launch_count(self):
        return self.L.orbx_launch_count(self.h)

    STAGES = ("pyramid", "fast_cells", "quadtree", "assemble", "blur", "orient_desc")

    def set_profiling(self, on):
        rc = self.L.orbx_set_profiling(self.h, int(bool(on)))
        if rc != OK:
            raise self._err(rc)

    def stage_ms(self, reset=False):
        ms = (C.c_double * 6)()
        calls = C.c_longlong(0)
        rc = self.L.orbx_get_stage_ms(self.h, ms, C.byref(calls))
        if rc != OK:
            raise self._err(rc)
        if reset:
            self.L.orbx_reset_stage_ms(self.h)
        return dict(zip(self.STAGES, list(ms))), calls.value

    # ---- mvImagePyramid (include/ORBextractor.h:83) and stage taps
    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        self.L.orbx_level_size(self.h, level, C.byref(w), C.byref(h))
        return w.value, h.value

    def mvImagePyramid(self, level, frame=0, padded=False):
        w, h = self.level_size(level)
        b = 38 if padded else 0
        out = np.zeros((h + b, w + b), np.uint8)
        rc = self.L.orbx_get_pyramid(self.h, frame, level, int(padded), _p(out), out.strides[0])
        if rc != OK:
            raise self._err(rc)
        return out

    def blurred(self, level, frame=0):
        w, h = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        rc = self.L.orbx_get_blurred(self.h, frame, level, _p(out), out.strides[0])
        if rc != OK:
            raise self._err(rc)
        return out

    def candidates(self, level, frame=0):
        n = self.L.orbx_get_candidates(self.h, frame, level, None, 0)
        if n < 0:
            raise self._err(n)
        out = np.zeros(max(n, 1), KP_DTYPE)
        self.L.orbx_get_candidates(self.h, frame, level, _p(out), n)
        return out[:n]

    def selected(self, level, frame=0):
        n = self.L.orbx_get_selected(self.h, frame, level, None, 0)
        if n < 0:
            raise self._err(n)
        out = np.zeros(max(n, 1), KP_DTYPE)
        self.L.orbx_get_selected(self.h, frame, level, _p(out), n)
        return out[:n]


class ORBmatcher:
    """Matching inner loops of `ORB_SLAM3::ORBmatcher` (include/ORBmatcher.h:36-103) and of
    `Frame::ComputeStereoFishEyeMatches` (src/Frame.cc:1060-1100) on one B200."""

    TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30  # src/ORBmatcher.cc:35-37

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self.L = lib()
        self.mfNNratio, self.mbCheckOrientation = np.float32(nnratio), bool(checkOri)
        self.device = device
        self.h = self.L.orbx_matcher_create(device)
        if not self.h:
            raise OrbxError(ERR_CUDA, self.L.orbx_matcher_last_error(None).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.orbx_matcher_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _chk(self, rc):
        if rc != OK:
            raise OrbxError(rc, self.L.orbx_matcher_last_error(self.h).decode())

    @staticmethod
    def DescriptorDistance(a, b) -> int:
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        assert a.size == 32 and b.size == 32
        return lib().orbx_descriptor_distance(_p(a), _p(b))

    def knnMatch(self, query, train):
        """BFMatcher(NORM_HAMMING).knnMatch(query, train, k=2) → (idx[nq,2], dist[nq,2])."""
        q = np.ascontiguousarray(query, np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
        idx = np.zeros((len(q), 2), np.int32)
        dist = np.zeros((len(q), 2), np.int32)
        self._chk(self.L.orbx_hamming_knn2(self.h, _p(q), len(q), _p(t), len(t), _p(idx), _p(dist)))
        return idx, dist

    def ratio_test(self, dist, ratio=0.7):
        dist = np.ascontiguousarray(dist, np.int32).reshape(-1, 2)
        keep = np.zeros(len(dist), np.uint8)
        self._chk(self.L.orbx_ratio_test(self.h, _p(dist), len(dist), float(ratio), _p(keep)))
        return keep.astype(bool)

    def top2_lists(self, query, train, cand, cand_off):
        q = np.ascontiguousarray(query, np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
        cand = np.ascontiguousarray(cand, np.int32)
        off = np.ascontiguousarray(cand_off, np.int32)
        assert len(off) == len(q) + 1
        bi, bd, si, sd = (np.zeros(len(q), np.int32) for _ in range(4))
        self._chk(self.L.orbx_hamming_top2_lists(self.h, _p(q), len(q), _p(t), len(t), _p(cand), _p(off), _p(bi), _p(bd), _p(si), _p(sd)))
        return bi, bd, si, sd

    def SearchByProjection(self, F, mp_proj5, mp_level, mp_flags, mp_obs, mp_desc, th=3.0, bFarPoints=False, thFarPoints=50.0):
        """`ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th, bFarPoints, thFarPoints)`
        (src/ORBmatcher.cc:43-213, Nleft == -1 frames).  F is a dict with the Frame members the function reads: `mvKeysUn`
        (keypoint records), `mDescriptors`, `bounds` = (mnMinX, mnMinY, mnMaxX, mnMaxY), `mvScaleFactors`, optional `mvuRight` and
        `kp_obs` (Observations() of the map point attached to each keypoint, -1 = none).  → (nmatches, assigned[n])."""
        kps = np.ascontiguousarray(F["mvKeysUn"], KP_DTYPE); desc = np.ascontiguousarray(F["mDescriptors"], np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(F["bounds"], np.float32); sf = np.ascontiguousarray(F["mvScaleFactors"], np.float32)
        ur = None if F.get("mvuRight") is None else np.ascontiguousarray(F["mvuRight"], np.float32)
        ko = None if F.get("kp_obs") is None else np.ascontiguousarray(F["kp_obs"], np.int32)
        p5 = np.ascontiguousarray(mp_proj5, np.float32).reshape(-1, 5); lv = np.ascontiguousarray(mp_level, np.int32)
        fl = np.ascontiguousarray(mp_flags, np.uint8); ob = np.ascontiguousarray(mp_obs, np.int32)
        md = np.ascontiguousarray(mp_desc, np.uint8).reshape(-1, 32)
        assert len(desc) == len(kps) and len(lv) == len(fl) == len(ob) == len(md) == len(p5)
        out = np.full(len(kps), -1, np.int32)
        n = C.c_int32(0)
        self._chk(self.L.orbx_search_by_projection(self.h, _p(kps), _p(desc), len(kps), _p(ur), _p(ko), _p(b), _p(sf), len(sf), _p(p5), _p(lv), _p(fl),
                                                   _p(ob), _p(md), len(p5), float(self.mfNNratio), float(th), int(bFarPoints), float(thFarPoints),
                                                   _p(out), C.byref(n)))
        return n.value, out

    def SearchByBoW(self, KF, F):
        """`ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches)` (src/ORBmatcher.cc:222-425,
        Nleft == -1 frames).  KF is a dict with `mDescriptors`, `angles` (mvKeysUn[i].angle), `map_points` (0 = null, 1 = good,
        2 = isBad()) and `mFeatVec` = (nodes, off, idx) CSR over ascending node ids; F a dict with `mDescriptors`, `angles`
        (mvKeys[i].angle) and `mFeatVec`.  → (nmatches, assigned[n_f]): the keyframe feature whose map point each frame feature got."""
        kd = np.ascontiguousarray(KF["mDescriptors"], np.uint8).reshape(-1, 32); ka = np.ascontiguousarray(KF["angles"], np.float32)
        km = np.ascontiguousarray(KF["map_points"], np.uint8)
        fd = np.ascontiguousarray(F["mDescriptors"], np.uint8).reshape(-1, 32); fa = np.ascontiguousarray(F["angles"], np.float32)
        kn, ko, ki = (np.ascontiguousarray(v, np.int32) for v in KF["mFeatVec"])
        fn, fo, fi = (np.ascontiguousarray(v, np.int32) for v in F["mFeatVec"])
        assert len(ka) == len(km) == len(kd) and len(fa) == len(fd) and len(ko) == len(kn) + 1 and len(fo) == len(fn) + 1
        out = np.full(len(fd), -1, np.int32)
        n = C.c_int32(0)
        self._chk(self.L.orbx_search_by_bow(self.h, _p(kd), _p(ka), len(kd), _p(km), _p(kn), _p(ko), _p(ki), len(kn), _p(fd), _p(fa), len(fd),
                                            _p(fn), _p(fo), _p(fi), len(fn), float(self.mfNNratio), int(self.mbCheckOrientation), _p(out), C.byref(n)))
        return n.value, out

    def SearchByBoWKeyFrames(self, KF1, KF2):
        """`ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12)` (src/ORBmatcher.cc:760-901).  Both sides
        are dicts like KF in SearchByBoW.  → (nmatches, matches12[n1]): the feature of keyframe 2 whose map point each feature of keyframe 1 got."""
        d1 = np.ascontiguousarray(KF1["mDescriptors"], np.uint8).reshape(-1, 32); a1 = np.ascontiguousarray(KF1["angles"], np.float32)
        m1 = np.ascontiguousarray(KF1["map_points"], np.uint8)
        d2 = np.ascontiguousarray(KF2["mDescriptors"], np.uint8).reshape(-1, 32); a2 = np.ascontiguousarray(KF2["angles"], np.float32)
        m2 = np.ascontiguousarray(KF2["map_points"], np.uint8)
        n1, o1, i1 = (np.ascontiguousarray(v, np.int32) for v in KF1["mFeatVec"])
        n2, o2, i2 = (np.ascontiguousarray(v, np.int32) for v in KF2["mFeatVec"])
        assert len(a1) == len(m1) == len(d1) and len(a2) == len(m2) == len(d2) and len(o1) == len(n1) + 1 and len(o2) == len(n2) + 1
        out = np.full(len(d1), -1, np.int32)
        n = C.c_int32(0)
        self._chk(self.L.orbx_search_by_bow_keyframes(self.h, _p(d1), _p(a1), len(d1), _p(m1), _p(n1), _p(o1), _p(i1), len(n1), _p(d2), _p(a2), len(d2),
                                                      _p(m2), _p(n2), _p(o2), _p(i2), len(n2), float(self.mfNNratio), int(self.mbCheckOrientation),
                                                      _p(out), C.byref(n)))
        return n.value, out

    def SearchForInitializationFrames(self, kps1, desc1, kps2, desc2, bounds, vbPrevMatched, windowSize=10):
        """`ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)` (src/ORBmatcher.cc:644-759), whole
        function incl. F2's grid query → (nmatches, vnMatches12, vbPrevMatched updated)."""
        k1 = np.ascontiguousarray(kps1, KP_DTYPE); k2 = np.ascontiguousarray(kps2, KP_DTYPE)
        d1 = np.ascontiguousarray(desc1, np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(desc2, np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(bounds, np.float32)
        prev = np.array(vbPrevMatched, np.float32).reshape(-1, 2).copy()
        assert len(prev) == len(k1) == len(d1) and len(k2) == len(d2)
        m12 = np.full(len(k1), -1, np.int32)
        n = C.c_int32(0)
        self._chk(self.L.orbx_search_for_initialization_frames(self.h, _p(k1), _p(d1), len(k1), _p(k2), _p(d2), len(k2), _p(b), _p(prev), int(windowSize),
                                                               float(self.mfNNratio), int(self.mbCheckOrientation), _p(m12), C.byref(n)))
        return n.value, m12, prev

    def rot_hist_filter(self, angle_a, angle_b):
        a = np.ascontiguousarray(angle_a, np.float32)
        b = np.ascontiguousarray(angle_b, np.float32)
        keep = np.zeros(len(a), np.uint8)
        self._chk(self.L.orbx_rot_hist_filter(self.h, _p(a), _p(b), len(a), _p(keep)))
        return keep.astype(bool)

    def SearchForInitialization(self, desc1, angle1, octave1, desc2, angle2, cand, cand_off):
        """ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:644-759) over explicit candidate lists
        → (nmatches, vnMatches12)."""
        d1 = np.ascontiguousarray(desc1, np.uint8).reshape(-1, 32)
        d2 = np.ascontiguousarray(desc2, np.uint8).reshape(-1, 32)
        a1 = np.ascontiguousarray(angle1, np.float32)
        a2 = np.ascontiguousarray(angle2, np.float32)
        o1 = np.ascontiguousarray(octave1, np.int32)
        cand = np.ascontiguousarray(cand, np.int32)
        off = np.ascontiguousarray(cand_off, np.int32)
        assert len(off) == len(d1) + 1 and len(a1) == len(d1) and len(a2) == len(d2)
        m12 = np.full(len(d1), -1, np.int32)
        n = C.c_int32(0)
        self._chk(self.L.orbx_search_for_initialization(self.h, _p(d1), _p(a1), _p(o1), len(d1), _p(d2), _p(a2), len(d2), _p(cand), _p(off),
                                                        float(self.mfNNratio), int(self.mbCheckOrientation), _p(m12), C.byref(n)))
        return n.value, m12

    def GetFeaturesInArea(self, keypoints_xy, octave, bounds, queries_xyr, minLevel=-1, maxLevel=-1):
        """Batched Frame::GetFeaturesInArea over the 64×48 grid (src/Frame.cc:387-418, :659-738) → (cand_off, cand)."""
        xy = np.ascontiguousarray(keypoints_xy, np.float32).reshape(-1, 2)
        oc = np.ascontiguousarray(octave, np.int32)
        q = np.ascontiguousarray(queries_xyr, np.float32).reshape(-1, 3)
        off = np.zeros(len(q) + 1, np.int32)
        tot = C.c_int32(0)
        b = [float(v) for v in bounds]
        self._chk(self.L.orbx_features_in_area(self.h, _p(xy), _p(oc), len(xy), b[0], b[1], b[2], b[3], _p(q), len(q), minLevel, maxLevel, _p(off), None, 0, C.byref(tot)))
        cand = np.zeros(max(tot.value, 1), np.int32)
        if tot.value:
            self._chk(self.L.orbx_features_in_area(self.h, _p(xy), _p(oc), len(xy), b[0], b[1], b[2], b[3], _p(q), len(q), minLevel, maxLevel, _p(off), _p(cand), tot.value, C.byref(tot)))
        return off, cand[: tot.value]

    def StereoTail(self, uL, uR, idx, dist, keep, mbf, mb):
        """Tail of Frame::ComputeStereoMatches (src/Frame.cc:862-914) over kNN+ratio matches → (n_kept, mvuRight, mvDepth)."""
        uL = np.ascontiguousarray(uL, np.float32); uR = np.ascontiguousarray(uR, np.float32)
        idx = np.ascontiguousarray(idx, np.int32); dist = np.ascontiguousarray(dist, np.int32)
        keep = np.ascontiguousarray(keep, np.uint8)
        ur = np.zeros(len(uL), np.float32); dp = np.zeros(len(uL), np.float32)
        n = C.c_int32(0)
        self._chk(self.L.orbx_stereo_tail(self.h, _p(uL), _p(uR), len(uL), len(uR), _p(idx), _p(dist), _p(keep), float(mbf), float(mb), _p(ur), _p(dp), C.byref(n)))
        return n.value, ur, dp

    # device-pointer forms (ints), asynchronous on stream()
    def UndistortKeyPoints(self, kps, K, distCoef):
        """`Frame::UndistortKeyPoints` (src/Frame.cc:749-782): keypoint records with `pt` undistorted (K = (fx, fy, cx, cy))."""
        kps = np.ascontiguousarray(kps, KP_DTYPE); D = np.ascontiguousarray(distCoef, np.float32)
        out = np.zeros_like(kps)
        self._chk(self.L.orbx_undistort_keypoints(self.h, _p(kps), len(kps), *[float(v) for v in K], _p(D), len(D), _p(out)))
        return out

    def undistort_points_device(self, d_xy, stride_in, n, K, distCoef, d_out, stride_out):
        D = np.ascontiguousarray(distCoef, np.float32)
        self._chk(self.L.orbx_undistort_points_device(self.h, d_xy, stride_in, n, *[float(v) for v in K], _p(D), len(D), d_out, stride_out))

    def ComputeImageBounds(self, cols, rows, K, distCoef):
        """`Frame::ComputeImageBounds` (src/Frame.cc:784-811) → (mnMinX, mnMaxX, mnMinY, mnMaxY)."""
        D = np.ascontiguousarray(distCoef, np.float32)
        b = np.zeros(4, np.float32)
        self._chk(self.L.orbx_image_bounds(self.h, int(cols), int(rows), *[float(v) for v in K], _p(D), len(D), _p(b)))
        return b

    def knn2_device(self, d_q, nq, d_db, ndb, idx_base, d_idx, d_dist):
        self._chk(self.L.orbx_hamming_knn2_device(self.h, d_q, nq, d_db, ndb, idx_base, d_idx, d_dist))

    def merge_device(self, d_idx_all, d_dist_all, n_shards, nq, d_idx, d_dist):
        self._chk(self.L.orbx_knn2_merge_device(self.h, d_idx_all, d_dist_all, n_shards, nq, d_idx, d_dist))

    def knn2_sharded(self, comm, d_q, nq, d_db_shard, ndb_shard, idx_base, d_idx, d_dist):
        """orbx_knn2_sharded: this rank's shard scan + ONE NCCL all-gather of the packed top-2 records + merge (device pointers as ints)."""
        rc = self.L.orbx_knn2_sharded(self.h, comm.h, d_q, nq, d_db_shard, ndb_shard, idx_base, d_idx, d_dist)
        if rc != OK:
            raise OrbxError(rc, self.L.orbx_comm_last_error(comm.h).decode())

    def tc_launches(self):
        """Brute-force kNN calls of this matcher that ran on the tensor-core kernel."""
        return self.L.orbx_debug_knn_tc_launches(self.h)

    def merge_packed_device(self, d_packed_all, n_shards, nq, d_idx, d_dist):
        self._chk(self.L.orbx_knn2_merge_packed_device(self.h, d_packed_all, n_shards, nq, d_idx, d_dist))

    def sync(self):
        self._chk(self.L.orbx_matcher_sync(self.h))

    def stream(self):
        return self.L.orbx_matcher_stream(self.h)


def ComputeStereoMatches(exL, exR, kpsL, descL, kpsR, descR, mbf, mb):
    """The classical `Frame::ComputeStereoMatches` (slot src/Frame.cc:813-915; restated upstream algorithm, see orbx.h) over the pyramids
    the two extractors hold from their last single-image calls → (n_stereo, mvuRight, mvDepth)."""
    L = lib()
    kL = np.ascontiguousarray(kpsL, KP_DTYPE); kR = np.ascontiguousarray(kpsR, KP_DTYPE)
    dL = np.ascontiguousarray(descL, np.uint8).reshape(-1, 32); dR = np.ascontiguousarray(descR, np.uint8).reshape(-1, 32)
    ur = np.zeros(len(kL), np.float32); dp = np.zeros(len(kL), np.float32)
    n = C.c_int32(0)
    rc = L.orbx_stereo_matches(exL.h, exR.h, _p(kL), _p(dL), len(kL), _p(kR), _p(dR), len(kR), float(mbf), float(mb), _p(ur), _p(dp), C.byref(n))
    if rc != OK:
        raise OrbxError(rc, L.orbx_last_error(exL.h).decode())
    return n.value, ur, dp


class Comm:
    """A rank of the DB-sharded kNN's communicator (orbx_comm: NCCL, bound at run time).  One process per GPU: rank 0 calls
    `Comm.unique_id()`, ships the 128 bytes to the others, every rank builds `Comm(world, rank, id, device)`.  One process with
    several GPUs: `Comm.create_all(devices)`."""

    def __init__(self, world=None, rank=None, uid=None, device=0, _handle=None):
        self.L = lib()
        if _handle is not None:
            self.h = _handle
            return
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(uid))
        self.h = self.L.orbx_comm_create(int(world), int(rank), buf, int(device))
        if not self.h:
            raise OrbxError(ERR_CUDA, self.L.orbx_comm_last_error(None).decode())

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = lib().orbx_comm_unique_id(buf)
        if rc != OK:
            raise OrbxError(rc, lib().orbx_comm_last_error(None).decode())
        return bytes(buf)

    @staticmethod
    def create_all(devices):
        L = lib()
        devs = (C.c_int * len(devices))(*devices)
        out = (C.c_void_p * len(devices))()
        rc = L.orbx_comm_create_all(out, len(devices), devs)
        if rc != OK:
            raise OrbxError(rc, L.orbx_comm_last_error(None).decode())
        return [Comm(_handle=out[i]) for i in range(len(devices))]

    @property
    def rank(self):
        return self.L.orbx_comm_rank(self.h)

    @property
    def world(self):
        return self.L.orbx_comm_world(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.orbx_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


def knn2_sharded_all(matchers, comms, d_query, nq, d_db, ndb, idx_base, d_idx, d_dist):
    """orbx_knn2_sharded_all: one thread drives every device of the process (lists indexed by rank; device pointers as ints)."""
    L = lib()
    n = len(matchers)
    arr = lambda vals: (C.c_void_p * n)(*vals)                     # noqa: E731
    i64 = lambda vals: (C.c_int64 * n)(*vals)                      # noqa: E731
    rc = L.orbx_knn2_sharded_all(arr([m.h for m in matchers]), arr([c.h for c in comms]), n, arr(d_query), int(nq), arr(d_db), i64(ndb), i64(idx_base),
                                 arr(d_idx), arr(d_dist))
    if rc != OK:
        msg = L.orbx_comm_last_error(None).decode() or "; ".join(L.orbx_comm_last_error(c.h).decode() for c in comms)
        raise OrbxError(rc, msg)


def extract_batch_multi(extractors, images, cap, vLappingArea=(0, 0)):
    """orbx_extract_batch_multi: `images` [B, H, W] uint8 split over the extractors' devices (contiguous slices, no collective)
    → (keypoints [B, cap], descriptors [B, cap, 32], n_out [B], mono_index [B])."""
    L = lib()
    images = np.ascontiguousarray(images, np.uint8)
    B, H, W = images.shape
    n = len(extractors)
    hs = (C.c_void_p * n)(*[e.h for e in extractors])
    ptrs = (C.c_void_p * B)(*[images.ctypes.data + b * H * W for b in range(B)])
    kps = np.zeros((B, cap), KP_DTYPE); desc = np.zeros((B, cap, 32), np.uint8)
    n_out = np.zeros(B, np.int32); mono = np.zeros(B, np.int32)
    rc = L.orbx_extract_batch_multi(hs, n, ptrs, B, H, W, W, None, 0, int(vLappingArea[0]), int(vLappingArea[1]), _p(kps), _p(desc), cap, _p(n_out), _p(mono))
    if rc != OK:
        raise OrbxError(rc, "; ".join(L.orbx_last_error(e.h).decode() for e in extractors))
    return kps, desc, n_out, mono


class ORBVocabulary:
    """`ORB_SLAM3::ORBVocabulary` = `DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>` (include/ORBVocabulary.h:29) for
    the calls the tracking front-end makes on it: `loadFromTextFile` and `transform(features, BowVector, FeatureVector,
    levelsup)` (src/Frame.cc:739-747).  The tree lives on one B200."""

    def __init__(self, device=0):
        self.L = lib()
        self.device = device
        self.h = None

    def _chk(self, rc):
        if rc != OK:
            raise OrbxError(rc, self.L.orbx_vocab_last_error(self.h).decode())

    def _adopt(self, h):
        self.close()
        if not h:
            raise OrbxError(ERR_ARG, self.L.orbx_vocab_last_error(None).decode())
        self.h = h

    def loadFromTextFile(self, path) -> bool:
        try:
            self._adopt(self.L.orbx_vocab_load_text(str(path).encode(), self.device))
        except OrbxError:
            return False      # the reference returns false on a malformed file
        return True

    def from_nodes(self, voc):
        """Node stream (dict as produced by dani_slam_b200.synth.vocabulary)."""
        par = np.ascontiguousarray(voc["parent"], np.int32); leaf = np.ascontiguousarray(voc["is_leaf"], np.uint8)
        desc = np.ascontiguousarray(voc["desc"], np.uint8); w = np.ascontiguousarray(voc["weight"], np.float64)
        self._adopt(self.L.orbx_vocab_create_from_nodes(_p(par), _p(leaf), _p(desc), _p(w), len(par), int(voc["k"]), int(voc["L"]),
                                                        int(voc["scoring"]), int(voc["weighting"]), self.device))
        return self

    def close(self):
        if getattr(self, "h", None):
            self.L.orbx_vocab_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def info(self):
        k, L_, nn, nw = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        self._chk(self.L.orbx_vocab_info(self.h, C.byref(k), C.byref(L_), C.byref(nn), C.byref(nw)))
        return dict(k=k.value, L=L_.value, n_nodes=nn.value, n_words=nw.value)

    def size(self):
        return self.info()["n_words"]

    def empty(self):
        return self.h is None or self.size() == 0

    def transform(self, descriptors, levelsup=4):
        """→ dict(word_id, node_id, bow_ids, bow_vals, fv_nodes, fv_off, fv_idx): BowVector and FeatureVector in map order."""
        desc = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        n = len(desc)
        m = max(n, 1)
        wid = np.zeros(m, np.uint32); nid = np.zeros(m, np.uint32)
        bi = np.zeros(m, np.uint32); bv = np.zeros(m, np.float64)
        fn = np.zeros(m, np.uint32); fo = np.zeros(m + 1, np.int32); fi = np.zeros(m, np.uint32)
        nb, nf = C.c_int(0), C.c_int(0)
        self._chk(self.L.orbx_bow_transform(self.h, _p(desc), n, int(levelsup), _p(wid), _p(nid), _p(bi), _p(bv), C.byref(nb), _p(fn), _p(fo),
                                            _p(fi), C.byref(nf)))
        nb, nf = nb.value, nf.value
        return dict(word_id=wid[:n], node_id=nid[:n], bow_ids=bi[:nb], bow_vals=bv[:nb], fv_nodes=fn[:nf], fv_off=fo[:nf + 1], fv_idx=fi[:fo[nf]])

    def transform_batch_device(self, d_desc, desc_stride, d_n, batch, cap, levelsup, d_word, d_node, d_bow_ids, d_bow_vals, d_n_bow, d_fv_nodes,
                               d_fv_off, d_fv_idx, d_n_fv):
        """Raw device pointers (ints); asynchronous on the vocabulary's stream."""
        self._chk(self.L.orbx_bow_transform_batch_device(self.h, d_desc, desc_stride, d_n, batch, cap, int(levelsup), d_word, d_node, d_bow_ids,
                                                         d_bow_vals, d_n_bow, d_fv_nodes, d_fv_off, d_fv_idx, d_n_fv))

    def sync(self):
        self._chk(self.L.orbx_vocab_sync(self.h))

    def stream(self):
        return self.L.orbx_vocab_stream(self.h)
