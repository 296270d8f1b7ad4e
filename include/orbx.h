/* orbx.h — C ABI of the B200-native ORB front-end (liborbx.so, hand-written sm_100a CUDA).
 *
 * This is the drop-in boundary for ONE hot path of DANI-SLAM: what the reference's
 * `ORB_SLAM3::ORBextractor` / `ORB_SLAM3::ORBmatcher` classes compute (file:line below are relative to
 * /root/reference).  The adapter headers include/ORBextractor.h and include/ORBmatcher.h keep the
 * reference's C++ signatures and forward to these entry points; INTEGRATION.md shows the binding.
 * Plain pointers and sizes only; no torch / OpenCV types.  No CPU fallback exists: every compute entry
 * point returns ORBX_ERR_CUDA when no usable device is present.
 *
 * Return convention: 0 = ok, -1 = empty image (the reference's `return -1`, src/ORBextractor.cc:1129),
 * other negative values = errors (see ORBX_ERR_*); the message is available via orbx_last_error().
 */
#ifndef ORBX_H
#define ORBX_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORBX_OK 0
#define ORBX_EMPTY (-1)         /* empty input image: reference returns -1 (ORBextractor.cc:1129-1130) */
#define ORBX_ERR_CAPACITY (-2)  /* caller's keypoint/descriptor capacity too small; n_out holds the need */
#define ORBX_ERR_GEOMETRY (-3)  /* image too small / too tall for the pyramid (reference would fault) */
#define ORBX_ERR_ARG (-4)       /* bad argument (null pointer, size beyond what the handle was made for) */
#define ORBX_ERR_CUDA (-5)      /* CUDA runtime error, or no device */

/* Layout-identical to cv::KeyPoint (7×4 bytes): pt.x, pt.y, size, angle, response, octave, class_id. */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orbx_keypoint;

typedef struct orbx_extractor orbx_extractor;

/* ---------------------------------------------------------------------------------------------
 * Extractor — replaces ORBextractor (include/ORBextractor.h:43-110, src/ORBextractor.cc).
 * ------------------------------------------------------------------------------------------- */

/* Number of CUDA devices visible (0 when none / driver missing). */
int orbx_device_count(void);

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 * (src/ORBextractor.cc:409-469) plus the device resources the handle owns: `device` ordinal, the
 * largest image (max_width × max_height) and the largest batch it will be asked to process.
 * Returns NULL on failure (orbx_last_error(NULL) tells why). */
orbx_extractor *orbx_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast,
                            int min_th_fast, int device, int max_width, int max_height, int max_batch);
void orbx_destroy(orbx_extractor *ex);

/* Last error text of this handle (or of the calling thread's last failed create when ex==NULL). */
const char *orbx_last_error(const orbx_extractor *ex);

/* GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares
 * (include/ORBextractor.h:64-78) and mnFeaturesPerLevel; arrays of nlevels entries, any may be NULL. */
int orbx_params(const orbx_extractor *ex, float *scale_factors, float *inv_scale_factors,
                float *level_sigma2, float *inv_level_sigma2, int32_t *features_per_level);

/* ORBextractor::operator()(image, mask, keypoints, descriptors, vLappingArea)
 * (src/ORBextractor.cc:1125-1207) for one HOST image (8-bit, 1 channel, `step` bytes per row).
 * rects_xywh = mvDynamicArea (include/ORBextractor.h:84) as n_rects×{x,y,w,h} in level-0 pixels;
 * lap0/lap1 = vLappingArea[0..1].  Writes *n_out keypoints (28-byte records) and *n_out×32 descriptor
 * bytes into caller buffers of capacity `cap` rows; *mono_index = the reference's return value.
 * This is the latency path (one call per frame, src/Frame.cc:420-427): the image is staged through page-locked memory owned
 * by the handle, all outputs return in one copy, the call synchronises once, and from the third call with the same (rows,
 * cols, cap) the copies and kernels replay as one CUDA graph (ORBX_NO_GRAPH in the environment keeps eager launches). */
int orbx_extract(orbx_extractor *ex, const uint8_t *image, int rows, int cols, size_t step,
                 const int32_t *rects_xywh, int n_rects, int lap0, int lap1, orbx_keypoint *keypoints,
                 uint8_t *descriptors, int cap, int *n_out, int *mono_index);

/* Frame-batch form of the same call (the throughput path): `batch` HOST images of identical size,
 * images[b] pointing at rows×cols bytes with `step`.  Outputs are batch×cap keypoints, batch×cap×32
 * descriptor bytes, n_out[batch], mono_index[batch].  Host↔device copies are inside the call.
 * Host buffers should be page-locked (orbx_host_alloc) for full PCIe rate; pageable works. */
int orbx_extract_batch(orbx_extractor *ex, const uint8_t *const *images, int batch, int rows, int cols,
                       size_t step, const int32_t *rects_xywh, int n_rects, int lap0, int lap1,
                       orbx_keypoint *keypoints, uint8_t *descriptors, int cap, int32_t *n_out,
                       int32_t *mono_index);

/* Measurement aid for the call above: moves exactly the bytes orbx_extract_batch moves (the frames in, cap-strided
 * keypoint and descriptor arrays out) over the same streams with the same chunking, and launches no kernel.  The outputs
 * are undefined.  bench.py reports the result as the transfer ceiling of the host path. */
int orbx_copy_only_batch(orbx_extractor *ex, const uint8_t *const *images, int batch, int rows, int cols, size_t step,
                         orbx_keypoint *keypoints, uint8_t *descriptors, int cap);

/* Device-resident batch: d_images = `batch` frames already in HBM (frame b at d_images + b*frame_stride,
 * rows of `step` bytes); outputs are DEVICE buffers of the same shapes as above.  Runs asynchronously on
 * the handle's stream; orbx_sync() waits.  Return code covers launch errors only; per-frame overflow
 * is reported by n_out[b] > cap after sync. */
int orbx_extract_batch_device(orbx_extractor *ex, const uint8_t *d_images, size_t frame_stride,
                              int batch, int rows, int cols, size_t step, const int32_t *rects_xywh,
                              int n_rects, int lap0, int lap1, orbx_keypoint *d_keypoints,
                              uint8_t *d_descriptors, int cap, int32_t *d_n_out, int32_t *d_mono_index);
int orbx_sync(orbx_extractor *ex);
/* The handle's cudaStream_t (as void*), so callers can record CUDA events on the launching stream. */
void *orbx_stream(orbx_extractor *ex);
/* Number of kernel launches issued by the handle since creation (bench `gpu_launches`). */
long long orbx_launch_count(const orbx_extractor *ex);

/* Per-stage device timing for bench.py's roofline: when on, CUDA events are recorded on the handle's
 * stream at the six stage boundaries of every batch call (pyramid, FAST cells, quadtree, assemble, blur,
 * orientation+rBRIEF); orbx_get_stage_ms returns the accumulated milliseconds per stage and the number of
 * batch calls they cover.  Off by default (the timed `value` region never runs with it on). */
int orbx_set_profiling(orbx_extractor *ex, int on);
int orbx_get_stage_ms(orbx_extractor *ex, double *ms6, long long *calls);
int orbx_reset_stage_ms(orbx_extractor *ex);

/* mvImagePyramid[level] (include/ORBextractor.h:83) of frame `frame` of the last call, copied to host.
 * padded!=0 → (w+38)×(h+38) plane with the REFLECT_101 border of src/ORBextractor.cc:1224-1230
 * (materialised lazily; the hot path never reads it); else the w×h level itself. */
int orbx_level_size(const orbx_extractor *ex, int level, int *width, int *height);
int orbx_get_pyramid(orbx_extractor *ex, int frame, int level, int padded, uint8_t *dst, size_t dst_step);

/* Stage taps of the last call, for stage-level parity tests: the 7×7 Gaussian-blurred level
 * (src/ORBextractor.cc:1171-1172), the per-level candidates entering DistributeOctTree (:909-916, after
 * the DANI filter) and the per-level selected keypoints (after :919-934).  Return the count. */
int orbx_get_blurred(orbx_extractor *ex, int frame, int level, uint8_t *dst, size_t dst_step);
int orbx_get_candidates(orbx_extractor *ex, int frame, int level, orbx_keypoint *out, int cap);
int orbx_get_selected(orbx_extractor *ex, int frame, int level, orbx_keypoint *out, int cap);

/* Classical rectified-stereo association — the slot of Frame::ComputeStereoMatches (src/Frame.cc:813-915).  This tree fills that
 * slot with a learned matcher (LightGlue, :822-860), which is out of scope; the entry point below is the published upstream
 * algorithm of the function (row bands of ±2·scale, Hamming search within one octave and the disparity range, 11×11 SAD
 * refinement over ±5 px with a parabola fit on the keypoint's pyramid level, disparity gate 0 <= d < mbf/mb, depth = mbf/d,
 * median cut at 1.5·1.4·median), restated in oracle/orb_oracle.cpp (orc_stereo_rowband; parity unpinned: no reference code to
 * run).  ex_left / ex_right are the two extractors right after their orbx_extract calls on the left / right image (their
 * pyramids stay on the device until the next call, like mvImagePyramid, src/ORBextractor.cc:1216); keypoints and
 * descriptors are what those calls returned.  mvu_right / mv_depth: n_left entries, -1 = no stereo.  HOST buffers. */
int orbx_stereo_matches(orbx_extractor *ex_left, orbx_extractor *ex_right, const orbx_keypoint *kps_left, const uint8_t *desc_left, int n_left,
                        const orbx_keypoint *kps_right, const uint8_t *desc_right, int n_right, float mbf, float mb, float *mvu_right, float *mv_depth,
                        int32_t *n_stereo);

/* Page-locked host memory helpers (cudaHostAlloc / cudaFreeHost). */
void *orbx_host_alloc(size_t bytes);
/* Same, write-combined (cudaHostAllocWriteCombined): for INPUT buffers the CPU only fills and the GPU only reads. */
void *orbx_host_alloc_wc(size_t bytes);
void orbx_host_free(void *p);

/* ---------------------------------------------------------------------------------------------
 * Matcher inner loops — replace ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2054-2070), the
 * best/second-best loops (:84-140, :273-325, :675-724, :812-864), ComputeThreeMaxima + the rotation
 * histogram (:2008-2049, :345-352) and Frame::BFmatcher.knnMatch(k=2) + Lowe ratio
 * (src/Frame.cc:45,1078-1085).  Descriptors are rows of 32 bytes.
 * ------------------------------------------------------------------------------------------- */
typedef struct orbx_matcher orbx_matcher;
/* Device context for matching (stream + scratch); cheap, but not free: adapters cache one per thread. */
orbx_matcher *orbx_matcher_create(int device);
void orbx_matcher_destroy(orbx_matcher *m);
const char *orbx_matcher_last_error(const orbx_matcher *m);
void *orbx_matcher_stream(orbx_matcher *m);
int orbx_matcher_sync(orbx_matcher *m);

/* BFMatcher(NORM_HAMMING).knnMatch(query, train, k=2): idx/dist are nq×2 (ascending distance, ties →
 * lower train index first); missing neighbours (ndb<2) are idx=-1, dist=INT32_MAX.  HOST buffers.
 * Large problems (nq >= 64, ndb >= 8192) run as an integer GEMM on the tensor cores (Hamming = |q| + |d| − 2·q·d over 0/1
 * bytes, tcgen05.mma.kind::i8, exact), small ones on the POPC kernel; both give identical results. */
int orbx_hamming_knn2(orbx_matcher *m, const uint8_t *query, int nq, const uint8_t *train, int64_t ndb,
                      int32_t *idx, int32_t *dist);
/* Same on DEVICE buffers, asynchronous on the matcher's stream.  idx_base is added to every train
 * index (global index of this shard's first row) so DB-sharded callers can merge shards directly. */
int orbx_hamming_knn2_device(orbx_matcher *m, const uint8_t *d_query, int nq, const uint8_t *d_train,
                             int64_t ndb, int64_t idx_base, int32_t *d_idx, int32_t *d_dist);
/* Merge per-shard top-2 lists (DEVICE, n_shards×nq×2, as produced by an all-gather of the above) into
 * the global top-2 by lexicographic (dist, idx) order — equals the unsharded result bit for bit. */
int orbx_knn2_merge_device(orbx_matcher *m, const int32_t *d_idx_all, const int32_t *d_dist_all,
                           int n_shards, int nq, int32_t *d_idx, int32_t *d_dist);
/* Same for PACKED per-shard records: shard g holds {idx[nq×2], dist[nq×2]} back to back (4·nq int32), which is what ONE
 * all-gather of the per-shard result moves when orbx_hamming_knn2_device wrote d_idx = rec, d_dist = rec + 2·nq. */
int orbx_knn2_merge_packed_device(orbx_matcher *m, const int32_t *d_packed_all, int n_shards, int nq, int32_t *d_idx, int32_t *d_dist);
/* Frame.cc:1085: keep[i] = (two neighbours) && (float)d0 < (float)d1 * ratio(double).  HOST / DEVICE buffers. */
int orbx_ratio_test(orbx_matcher *m, const int32_t *dist, int nq, double ratio, uint8_t *keep);
int orbx_ratio_test_device(orbx_matcher *m, const int32_t *d_dist, int nq, double ratio, uint8_t *d_keep);

/* Best / second-best over explicit candidate lists (the a12 loops, src/ORBmatcher.cc:84-121 and its siblings): query i is
 * compared with train rows cand[cand_off[i] .. cand_off[i+1]); strict '<' (first candidate wins ties); the loop starts from
 * 256 / 256 like the reference, so best_dist / second_dist default to 256 and a distance of 256 is never recorded.
 * second_idx (may be NULL) is the train row of the second best (-1: none) — the keypoint whose octave the reference keeps as
 * bestLevel2 for the level-aware ratio rule of :123-128.  HOST buffers. */
int orbx_hamming_top2_lists(orbx_matcher *m, const uint8_t *query, int nq, const uint8_t *train, int64_t ndb,
                            const int32_t *cand, const int32_t *cand_off, int32_t *best_idx, int32_t *best_dist,
                            int32_t *second_idx, int32_t *second_dist);
/* ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th, bFarPoints, thFarPoints)
 * (src/ORBmatcher.cc:43-213) for frames with Nleft == -1 (monocular, rectified stereo, RGB-D), whole function: the window
 * query F.GetFeaturesInArea(projX, projY, r·scale[level], level-1, level), the occupied-keypoint and right-image skips
 * (:88-97), the level-aware best / second-best loop and ratio rule (:101-128), and the in-order attachment of map points to
 * keypoints (later map points see earlier attachments).
 *   frame:      keypoints_un = F.mvKeysUn (n records), descriptors n×32, u_right = F.mvuRight (NULL: all -1), kp_obs[i] =
 *               F.mvpMapPoints[i]->Observations(), -1 for a null pointer (NULL: all -1), bounds4 = mnMinX, mnMinY, mnMaxX,
 *               mnMaxY, scale_factors = F.mvScaleFactors (n_levels entries)
 *   map points: mp_proj5 = n_mp × {mTrackProjX, mTrackProjY, mTrackProjXR, mTrackViewCos, mTrackDepth}, mp_level =
 *               mnTrackScaleLevel, mp_flags bit 0 = mbTrackInView, bit 1 = isBad(), mp_obs = Observations(), mp_desc =
 *               GetDescriptor() rows
 *   out:        assigned[i] = index of the map point the call wrote into F.mvpMapPoints[i], or -1; *n_matches = return value
 * HOST buffers. */
int orbx_search_by_projection(orbx_matcher *m, const orbx_keypoint *keypoints_un, const uint8_t *descriptors, int n, const float *u_right,
                              const int32_t *kp_obs, const float *bounds4, const float *scale_factors, int n_levels, const float *mp_proj5,
                              const int32_t *mp_level, const uint8_t *mp_flags, const int32_t *mp_obs, const uint8_t *mp_desc, int n_mp,
                              float nnratio, float th, int far_points, float th_far, int32_t *assigned, int32_t *n_matches);
/* ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches) (src/ORBmatcher.cc:222-425), frames with
 * Nleft == -1, whole function: the walk over the vocabulary nodes both feature vectors share, per keyframe feature with a good map
 * point the best / second-best scan over the frame's features of the same node that hold no match yet (ordered: an earlier match
 * removes its frame feature from every later scan, :281-282), TH_LOW, the fp32 ratio test, and the rotation-histogram purge.
 *   keyframe:  kf_desc n_kf×32 = pKF->mDescriptors, kf_angle = pKF->mvKeysUn[i].angle, kf_mp[i] = 0: GetMapPointMatches()[i] is null,
 *              1: a good map point, 2: isBad(); feature vector pKF->mFeatVec as CSR: kf_nodes[kf_nn] ascending node ids, kf_off[kf_nn+1],
 *              kf_idx[] the feature indices of each node in their stored order
 *   frame:     f_desc n_f×32 = F.mDescriptors, f_angle = F.mvKeys[i].angle, F.mFeatVec as CSR (f_nodes, f_off, f_idx)
 *   out:       assigned[i] = keyframe feature whose map point the call leaves in vpMapPointMatches[i], or -1; *n_matches = return value
 * HOST buffers. */
int orbx_search_by_bow(orbx_matcher *m, const uint8_t *kf_desc, const float *kf_angle, int n_kf, const uint8_t *kf_mp, const int32_t *kf_nodes,
                       const int32_t *kf_off, const int32_t *kf_idx, int kf_nn, const uint8_t *f_desc, const float *f_angle, int n_f,
                       const int32_t *f_nodes, const int32_t *f_off, const int32_t *f_idx, int f_nn, float nnratio, int check_orientation,
                       int32_t *assigned, int32_t *n_matches);
/* ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12) (src/ORBmatcher.cc:760-901), whole function:
 * as above between two keyframes — a candidate of keyframe 2 needs a good map point of its own and leaves the scans once matched
 * (vbMatched2), the threshold is bestDist1 < TH_LOW.  mp1 / mp2 as kf_mp, feature vectors as CSR.
 *   out: matches12[i] = feature of keyframe 2 whose map point the call leaves in vpMatches12[i], or -1; *n_matches = return value
 * HOST buffers. */
int orbx_search_by_bow_keyframes(orbx_matcher *m, const uint8_t *desc1, const float *angle1, int n1, const uint8_t *mp1, const int32_t *nodes1,
                                 const int32_t *off1, const int32_t *idx1, int nn1, const uint8_t *desc2, const float *angle2, int n2,
                                 const uint8_t *mp2, const int32_t *nodes2, const int32_t *off2, const int32_t *idx2, int nn2, float nnratio,
                                 int check_orientation, int32_t *matches12, int32_t *n_matches);
/* Rotation-consistency filter: bin = round((a-b [+360]) / 30) (quirk Q10), keep matches in the three
 * fullest bins subject to the 0.1·max rule.  HOST / DEVICE buffers; n matches. */
int orbx_rot_hist_filter(orbx_matcher *m, const float *angle_a, const float *angle_b, int n, uint8_t *keep);
int orbx_rot_hist_filter_device(orbx_matcher *m, const float *d_angle_a, const float *d_angle_b, int n, uint8_t *d_keep);
/* ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:644-759) with the per-query candidate lists
 * (the output of F2.GetFeaturesInArea(vbPrevMatched[i1], windowSize, 0, 0), src/ORBmatcher.cc:666) given
 * explicitly: query order, the vMatchedDistance lock, match stealing, TH_LOW, the fp32 ratio test and the
 * rotation-histogram purge are replayed exactly.  matches12[n1] = vnMatches12; *n_matches = return value.
 * HOST buffers. */
int orbx_search_for_initialization(orbx_matcher *m, const uint8_t *desc1, const float *angle1, const int32_t *octave1,
                                   int n1, const uint8_t *desc2, const float *angle2, int n2, const int32_t *cand,
                                   const int32_t *cand_off, float nnratio, int check_orientation, int32_t *matches12,
                                   int32_t *n_matches);
/* ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, vbPrevMatched, vnMatches12, windowSize) (src/ORBmatcher.cc:644-759),
 * whole function: F2's 64×48 grid, the window query per level-0 keypoint of F1, the ordered replay above and the vbPrevMatched
 * update of :754-756.  kps1 / kps2 = mvKeysUn of both frames, bounds4 = mnMinX, mnMinY, mnMaxX, mnMaxY, prev_matched_xy =
 * vbPrevMatched (n1×2, updated in place).  HOST buffers. */
int orbx_search_for_initialization_frames(orbx_matcher *m, const orbx_keypoint *kps1, const uint8_t *desc1, int n1, const orbx_keypoint *kps2,
                                          const uint8_t *desc2, int n2, const float *bounds4, float *prev_matched_xy, int window_size,
                                          float nnratio, int check_orientation, int32_t *matches12, int32_t *n_matches);
/* Frame::AssignFeaturesToGrid + Frame::GetFeaturesInArea (src/Frame.cc:387-418, :659-738; 64×48 grid,
 * include/Frame.h:52-53) for a batch of queries: keypoints_xy = n×{x,y} (mvKeysUn), octave[n], image bounds
 * (mnMinX, mnMinY, mnMaxX, mnMaxY), queries = nq×{x, y, r}.  cand_off[nq+1] and cand[*total_out] are the
 * vIndices lists in the reference's order (cell column, cell row, insertion order).  Call with cand=NULL to
 * size the output.  HOST buffers. */
int orbx_features_in_area(orbx_matcher *m, const float *keypoints_xy, const int32_t *octave, int n, float min_x, float min_y,
                          float max_x, float max_y, const float *queries_xyr, int nq, int min_level, int max_level,
                          int32_t *cand_off, int32_t *cand, int cap, int32_t *total_out);
/* Stereo association tail of Frame::ComputeStereoMatches (src/Frame.cc:862-914) fed by the Hamming kNN + Lowe
 * ratio of src/Frame.cc:1078-1085: for left keypoint i with keep[i], iR = idx[2i], distance = dist[2i]; disparity
 * gate 0 <= uL-uR < mbf/mb, depth = mbf/disparity, then the 1.5·median distance cut.  mvu_right / mv_depth are
 * n_left arrays (-1 = no stereo); *n_kept = stereo points that survive.  HOST buffers. */
int orbx_stereo_tail(orbx_matcher *m, const float *u_left, const float *u_right, int n_left, int n_right, const int32_t *idx,
                     const int32_t *dist, const uint8_t *keep, float mbf, float mb, float *mvu_right, float *mv_depth,
                     int32_t *n_kept);
/* Frame::UndistortKeyPoints (src/Frame.cc:749-782): kps_un = kps with pt replaced by cv::undistortPoints(pt, K, D, I, K);
 * a plain copy when D[0] == 0 (:751).  dist_coef = k1 k2 p1 p2 [k3 …] (n_coef <= 14).  HOST buffers; bit-identical to
 * cv2's result. */
int orbx_undistort_keypoints(orbx_matcher *m, const orbx_keypoint *kps, int n, float fx, float fy, float cx, float cy, const float *dist_coef,
                             int n_coef, orbx_keypoint *kps_un);
/* Same on DEVICE (x, y) pairs that are stride_in / stride_out floats apart (2 = packed pairs, 7 = keypoint records, e.g.
 * the array orbx_extract_batch_device wrote); asynchronous on the matcher's stream; in place is allowed. */
int orbx_undistort_points_device(orbx_matcher *m, const float *d_xy, int stride_in, int n, float fx, float fy, float cx, float cy,
                                 const float *dist_coef, int n_coef, float *d_out, int stride_out);
/* Frame::ComputeImageBounds (src/Frame.cc:784-811): bounds4 = mnMinX, mnMaxX, mnMinY, mnMaxY. */
int orbx_image_bounds(orbx_matcher *m, int cols, int rows, float fx, float fy, float cx, float cy, const float *dist_coef, int n_coef,
                      float *bounds4);
/* ORBmatcher::DescriptorDistance for one pair on the host (inline popcount; no device involved). */
int orbx_descriptor_distance(const uint8_t *a, const uint8_t *b);

/* ---------------------------------------------------------------------------------------------
 * Bag-of-words transform of ORB descriptors.  Replaces, for this step, the reference's
 * vendored DBoW2 as Frame::ComputeBoW / KeyFrame::ComputeBoW drive it (src/Frame.cc:739-747, src/KeyFrame.cc:
 * mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4)):
 *   TemplatedVocabulary::loadFromTextFile  Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1338-1423
 *   TemplatedVocabulary::transform         :1127-1194 (image) and :1218-1262 (one descriptor down the tree)
 *   BowVector::addWeight/normalize         Thirdparty/DBoW2/DBoW2/BowVector.cpp:32-85
 *   FeatureVector::addFeature              Thirdparty/DBoW2/DBoW2/FeatureVector.cpp:30-46
 * Scoring / weighting codes are DBoW2's (BowVector.h:39-56): scoring 0 L1_NORM … 5 DOT_PRODUCT, weighting 0 TF_IDF,
 * 1 TF, 2 IDF, 3 BINARY.  Results are bit-identical to DBoW2's, doubles included.
 * ------------------------------------------------------------------------------------------- */
typedef struct orbx_vocab orbx_vocab;
/* Vocabulary from a node stream in the order of the reference's text format: node i+1 (node 0 is the root) has
 * parent[i], a leaf flag (leaves get word ids in stream order), a 32-byte descriptor and a weight.  The tree is
 * uploaded to `device` once and stays resident.  NULL on error (orbx_vocab_last_error(NULL)). */
orbx_vocab *orbx_vocab_create_from_nodes(const int32_t *parent, const uint8_t *is_leaf, const uint8_t *desc, const double *weight,
                                         int n_nodes, int k, int L, int scoring, int weighting, int device);
/* Same from a file in the ORBvoc.txt format ("k L scoring weighting", then "parent isLeaf d0 … d31 weight" per
 * node).  Blank lines are ignored (the reference's eof loop reads a phantom node from a trailing one). */
orbx_vocab *orbx_vocab_load_text(const char *path, int device);
void orbx_vocab_destroy(orbx_vocab *v);
const char *orbx_vocab_last_error(const orbx_vocab *v);
int orbx_vocab_info(const orbx_vocab *v, int *k, int *L, int *n_nodes, int *n_words);
void *orbx_vocab_stream(orbx_vocab *v);
int orbx_vocab_sync(orbx_vocab *v);
/* transform(features, BowVector, FeatureVector, levelsup) for one image, HOST buffers.  word_id / node_id (n each,
 * may be NULL) are the per-feature word and the ancestor at level L - levelsup (0 = root when that level is <= 0 or
 * below the leaf, where the reference reads an uninitialised value).  The BowVector comes back as (bow_ids,
 * bow_vals) and the FeatureVector as CSR (fv_nodes, fv_off, fv_idx), both in std::map order; arrays hold n entries
 * (fv_off n+1).  Calls on one handle serialise (the reference shares one const vocabulary between its threads). */
int orbx_bow_transform(orbx_vocab *v, const uint8_t *desc, int n, int levelsup, uint32_t *word_id, uint32_t *node_id, uint32_t *bow_ids,
                       double *bow_vals, int32_t *n_bow, uint32_t *fv_nodes, int32_t *fv_off, uint32_t *fv_idx, int32_t *n_fv);
/* Batched form on DEVICE buffers, asynchronous on the vocabulary's stream: image b has d_n[b] (<= cap) descriptors
 * at d_desc + b*desc_stride_bytes — the layout orbx_extract_batch_device writes — and its outputs start at row
 * b*cap of every array (d_fv_off: b*(cap+1)). */
int orbx_bow_transform_batch_device(orbx_vocab *v, const uint8_t *d_desc, size_t desc_stride_bytes, const int32_t *d_n, int batch, int cap,
                                    int levelsup, uint32_t *d_word_id, uint32_t *d_node_id, uint32_t *d_bow_ids, double *d_bow_vals,
                                    int32_t *d_n_bow, uint32_t *d_fv_nodes, int32_t *d_fv_off, uint32_t *d_fv_idx, int32_t *d_n_fv);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU entry points.  Work shards with no data-path collective for frame batches, and with ONE
 * all-gather of the per-shard top-2 records (NCCL over NVLink / NVSwitch) for the DB-sharded Hamming kNN of BASELINE config 4.
 * NCCL is bound at run time (libnccl.so.2, or the path in ORBX_NCCL_LIB); hosts that never call these need no NCCL.
 * ------------------------------------------------------------------------------------------- */
typedef struct orbx_comm orbx_comm;
/* One process per GPU: rank 0 obtains the 128-byte id and distributes it by any means, every rank then creates its
 * communicator on its own device.  NULL / negative on failure (orbx_comm_last_error(NULL)). */
int orbx_comm_unique_id(uint8_t id128[128]);
orbx_comm *orbx_comm_create(int world, int rank, const uint8_t id128[128], int device);
/* One process driving n_devices GPUs (the C++ SLAM host): comms[i] lives on devices[i], rank i of n_devices. */
int orbx_comm_create_all(orbx_comm **comms, int n_devices, const int *devices);
void orbx_comm_destroy(orbx_comm *c);
const char *orbx_comm_last_error(const orbx_comm *c);
int orbx_comm_rank(const orbx_comm *c);
int orbx_comm_world(const orbx_comm *c);
/* BFMatcher.knnMatch(k=2) against a database split into contiguous index ranges over the ranks: this rank scans its shard
 * (rows idx_base … idx_base+ndb_shard-1 of the whole database), ONE ncclAllGather moves every rank's packed top-2 record
 * {idx[nq×2], dist[nq×2]} (32 KB at nq = 2000), and every rank merges world×2 candidates per query by (dist, global index) —
 * bit-identical to the unsharded search, ties included.  DEVICE buffers on the matcher's device; asynchronous on the
 * matcher's stream.  Collective: every rank of the communicator must call it with the same nq. */
int orbx_knn2_sharded(orbx_matcher *m, orbx_comm *c, const uint8_t *d_query, int nq, const uint8_t *d_db_shard, int64_t ndb_shard, int64_t idx_base,
                      int32_t *d_idx, int32_t *d_dist);
/* Same for one thread that owns all n matchers / communicators of a process (arrays indexed by rank): the local scans are
 * enqueued on every device, the n all-gathers form one NCCL group, then every device merges. */
int orbx_knn2_sharded_all(orbx_matcher *const *ms, orbx_comm *const *cs, int n, const uint8_t *const *d_query, int nq, const uint8_t *const *d_db_shard,
                          const int64_t *ndb_shard, const int64_t *idx_base, int32_t *const *d_idx, int32_t *const *d_dist);
/* orbx_extract_batch over several devices: handle i (created on its own device) processes the contiguous slice
 * [i·batch/n, (i+1)·batch/n) of the frames on its own host thread; no collective.  Arguments as orbx_extract_batch. */
int orbx_extract_batch_multi(orbx_extractor *const *exs, int n_handles, const uint8_t *const *images, int batch, int rows, int cols, size_t step,
                             const int32_t *rects_xywh, int n_rects, int lap0, int lap1, orbx_keypoint *keypoints, uint8_t *descriptors, int cap,
                             int32_t *n_out, int32_t *mono_index);

/* Test hook: how many brute-force kNN calls of this matcher ran on the tensor-core kernel (csrc/orbx_knn_tc.cu: tcgen05.mma.kind::i8 over 0/1
 * byte operands, exact int32 accumulation, fused top-2 epilogue; used for nq >= 64 and ndb >= 8192; ORBX_KNN_POPC in the environment
 * keeps the POPC kernel — same results). */
long long orbx_debug_knn_tc_launches(const orbx_matcher *m);

/* Test hook: number of (frame, level) pairs of the last batch call that the histogram quadtree kernel handed to
 * the general quadtree kernel (trees deeper than its table); -1 if the histogram kernel is disabled. */
int orbx_debug_deep_count(orbx_extractor *ex);

/* Test hook: number of FAST cells of the last batch call whose corner candidates did not fit the two-phase kernel's queue
 * and were redone by the single-phase kernel; -1 when ORBX_FAST_V1 routes everything through the latter. */
int orbx_debug_dense_count(orbx_extractor *ex);

/* Test hook: the device/host port of libstdc++ std::sort used by the quadtree (see stdsort_port.h),
 * run on the host; perm_out = resulting order of original indices. */
void orbx_debug_sort_nodes(const int32_t *sizes, const int32_t *ulx, int n, int32_t *perm_out);
/* Test hooks that run single device kernels on arrays (return 0 / ORBX_ERR_CUDA). */
int orbx_debug_sort_nodes_device(int device, const int32_t *sizes, const int32_t *ulx, int n, int32_t *perm_out);
int orbx_debug_sincos_device(int device, const float *angles, int n, float *sin_out, float *cos_out);
int orbx_debug_atan2_device(int device, const float *y, const float *x, int n, float *deg_out);

#ifdef __cplusplus
}
#endif
#endif
