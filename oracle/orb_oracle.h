/* orb_oracle.h — C ABI of the CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT).
 *
 * The oracle restates, on the CPU, the algorithm of the reference's ORB front-end
 * (/root/reference/src/ORBextractor.cc, src/ORBmatcher.cc inner loops, src/Frame.cc:1060-1100)
 * over scalar restatements of the OpenCV primitives that reference calls.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * it; nothing under dani_slam_b200/ links, imports or executes it.
 *
 * Parity pin: the reference ships no golden vectors for this path.  The oracle is
 * pinned two ways: (1) every primitive and the whole pipeline are checked bit-for-bit against the
 * container's real OpenCV (cv2 4.13.0) through oracle/cv2_oracle.py and the committed fixtures in
 * tests/golden/; (2) oracle/_ref compiles the reference's own src/ORBextractor.cc, unmodified, against
 * a minimal opencv2 shim backed by these primitives, and tests compare the two end to end.
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Layout-identical to cv::KeyPoint (28 bytes). */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orc_keypoint;

typedef struct orc_extractor orc_extractor;

/* ---- extractor (ORBextractor.cc:409-469 ctor, :1125-1207 operator()) ---- */
orc_extractor *orc_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th);
void orc_destroy(orc_extractor *ex);
/* arrays of nlevels entries each (any may be NULL) */
void orc_params(const orc_extractor *ex, float *sf, float *inv_sf, float *sigma2, float *inv_sigma2,
                int *quota, int *umax16);
/* returns 0 ok, -1 empty image (reference :1129), -2 capacity too small. */
int orc_extract(orc_extractor *ex, const uint8_t *img, int rows, int cols, size_t step,
                const int32_t *rects_xywh, int n_rects, int lap0, int lap1, orc_keypoint *kps,
                uint8_t *desc, int cap, int *n_out, int *mono_index);
/* stage taps of the LAST orc_extract call, for stage-level parity tests */
int orc_level_size(const orc_extractor *ex, int level, int *w, int *h);
int orc_get_level(const orc_extractor *ex, int level, int padded, uint8_t *dst, size_t dst_step);
int orc_get_blurred(const orc_extractor *ex, int level, uint8_t *dst, size_t dst_step);
int orc_get_candidates(const orc_extractor *ex, int level, orc_keypoint *out, int cap);
int orc_get_selected(const orc_extractor *ex, int level, orc_keypoint *out, int cap);

/* ---- primitives, exposed for known-answer tests against cv2 ---- */
void orc_resize_linear_u8(const uint8_t *src, int sw, int sh, size_t sstep, uint8_t *dst, int dw,
                          int dh, size_t dstep);
void orc_border_reflect101_u8(const uint8_t *src, int w, int h, size_t sstep, uint8_t *dst,
                              size_t dstep, int border);
void orc_gaussian7_u8(const uint8_t *src, int w, int h, size_t sstep, uint8_t *dst, size_t dstep);
/* FAST-9_16 with NMS on a w×h ROI; writes (x,y,score) triples; returns the count (may exceed cap) */
int orc_fast9_nms(const uint8_t *roi, int w, int h, size_t step, int threshold, int32_t *xys, int cap);
float orc_fast_atan2(float y, float x);
int orc_cvround(float v);
void orc_sincos(float angle_rad, float *s, float *c);
void orc_sincos_array(const float *a, int64_t n, float *s, float *c, int nthreads);
/* quadtree on an explicit candidate list (ORBextractor.cc:555-779); returns number kept */
int orc_distribute(const orc_keypoint *in, int n, int minX, int maxX, int minY, int maxY, int N,
                   orc_keypoint *out, int cap);
/* std::sort with the reference's compareNodes on (size, ULx) pairs; perm_out = resulting order of
 * the original indices (the oracle's definition of tie order). */
void orc_sort_nodes(const int32_t *sizes, const int32_t *ulx, int n, int32_t *perm_out);

/* ---- matcher inner loops ---- */
int orc_descriptor_distance(const uint8_t *a, const uint8_t *b); /* ORBmatcher.cc:2054-2070 */
/* BFMatcher(NORM_HAMMING).knnMatch(k=2): idx/dist are nq×2; missing entries = -1 / INT32_MAX */
void orc_knn2(const uint8_t *q, int nq, const uint8_t *db, int64_t ndb, int32_t *idx, int32_t *dist,
              int nthreads);
/* Frame.cc:1085 Lowe test: keep[i] = has two && d0 < d1*0.7 (float*double) */
void orc_ratio_test(const int32_t *dist, int nq, double ratio, uint8_t *keep);
/* top-2 over explicit candidate lists (ORBmatcher.cc:84-140 style): best, second distance, best idx */
void orc_top2_lists(const uint8_t *q, int nq, const uint8_t *db, const int32_t *cand,
                    const int32_t *cand_off, int32_t *best_idx, int32_t *best_dist,
                    int32_t *second_idx /* may be NULL */, int32_t *second_dist);
/* rotation histogram + ComputeThreeMaxima filter (ORBmatcher.cc:345-352, :2008-2049, :725-748) */
void orc_three_maxima(const int32_t *counts, int L, int32_t *ind3); /* ORBmatcher.cc:2008-2049 on list sizes */
void orc_rot_hist_filter(const float *angle_a, const float *angle_b, int n, uint8_t *keep);
/* SearchForInitialization core (ORBmatcher.cc:644-759) with explicit candidate lists. */
int orc_search_init(const uint8_t *d1, const float *ang1, const int32_t *oct1, int n1,
                    const uint8_t *d2, const float *ang2, int n2, const int32_t *cand,
                    const int32_t *cand_off, float nnratio, int check_ori, int32_t *matches12);

/* Frame grid: AssignFeaturesToGrid + GetFeaturesInArea (src/Frame.cc:387-418, :659-738); returns total candidates */
int orc_features_in_area(const float *xy, const int32_t *octave, int n, float minX, float minY, float maxX, float maxY,
                         const float *queries, int nq, int minLevel, int maxLevel, int32_t *cand_off, int32_t *cand, int cap);
/* stereo association tail (src/Frame.cc:862-914) over kNN+ratio matches; returns number of stereo points kept */
/* SearchByProjection(Frame&, vector<MapPoint*>&, …) for Nleft == -1 frames (ORBmatcher.cc:43-213) */
int orc_search_by_projection(const float *xy, const int32_t *octave, const uint8_t *desc, int n, const float *u_right, const int32_t *kp_obs,
                             float minX, float minY, float maxX, float maxY, const float *scale_factors, const float *mp_proj5,
                             const int32_t *mp_level, const uint8_t *mp_flags, const int32_t *mp_obs, const uint8_t *mp_desc, int m,
                             float nnratio, float th, int far_points, float th_far, int32_t *assigned);
/* SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) for Nleft == -1 frames (ORBmatcher.cc:222-425); feature vectors as CSR */
int orc_search_by_bow(const uint8_t *kf_desc, const float *kf_angle, int n_kf, const uint8_t *kf_mp, const int32_t *kf_nodes, const int32_t *kf_off,
                      const int32_t *kf_idx, int kf_nn, const uint8_t *f_desc, const float *f_angle, int n_f, const int32_t *f_nodes,
                      const int32_t *f_off, const int32_t *f_idx, int f_nn, float nnratio, int check_ori, int32_t *assigned);
/* SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&) (ORBmatcher.cc:760-901) */
int orc_search_by_bow_kf(const uint8_t *desc1, const float *angle1, int n1, const uint8_t *mp1, const int32_t *nodes1, const int32_t *off1,
                         const int32_t *idx1, int nn1, const uint8_t *desc2, const float *angle2, int n2, const uint8_t *mp2, const int32_t *nodes2,
                         const int32_t *off2, const int32_t *idx2, int nn2, float nnratio, int check_ori, int32_t *matches12);
/* classical rectified-stereo association over the two extractors' pyramids (restated upstream algorithm; parity unpinned) */
int orc_stereo_rowband(const orc_extractor *exL, const orc_extractor *exR, const orc_keypoint *kL, const uint8_t *dL, int nL,
                       const orc_keypoint *kR, const uint8_t *dR, int nR, float mbf, float mb, float *uRight, float *depth);
int orc_stereo_tail(const float *uL, const float *uR, int nL, int nR, const int32_t *idx, const int32_t *dist,
                    const uint8_t *keep, float mbf, float mb, float *uRight, float *depth);

/* Frame::UndistortKeyPoints / ComputeImageBounds (src/Frame.cc:749-811): cv::undistortPoints(pts, K, D, R = I, P = K) on n (x, y)
 * float pairs; D has nd = 4 or 5 (or up to 14) coefficients; returns immediately with a copy when D[0] == 0 (:751).
 * bounds[4] = mnMinX, mnMaxX, mnMinY, mnMaxY from the four image corners. */
void orc_undistort_points(const float *xy, int n, float fx, float fy, float cx, float cy, const float *D, int nd, float *out);
void orc_image_bounds(int cols, int rows, float fx, float fy, float cx, float cy, const float *D, int nd, float *bounds);

/* ---- bag of words (bow_oracle.cpp): DBoW2 TemplatedVocabulary::transform as Frame::ComputeBoW calls it ---- */
/* node stream like loadFromTextFile (TemplatedVocabulary.h:1378-1420): node i+1 has parent[i] (0 = root), leaf flag, 32-byte descriptor, weight */
void *orc_vocab_from_nodes(const int32_t *parent, const uint8_t *is_leaf, const uint8_t *desc, const double *weight, int n_nodes,
                           int k, int L, int scoring, int weighting);
void *orc_vocab_load_text(const char *path);
void orc_vocab_free(void *v);
int orc_vocab_words(void *v);
int orc_vocab_nodes(void *v);
/* per feature word/node id; BowVector as (ids, vals) and FeatureVector as CSR, both in std::map order */
int orc_bow_transform(void *v, const uint8_t *desc, int n, int levelsup, uint32_t *word_id, uint32_t *node_id, uint32_t *bow_ids,
                      double *bow_vals, int *n_bow, uint32_t *fv_nodes, int32_t *fv_off, uint32_t *fv_idx, int *n_fv);
double orc_bow_score_l1(const uint32_t *ids1, const double *v1, int n1, const uint32_t *ids2, const double *v2, int n2);

#ifdef __cplusplus
}
#endif
#endif
