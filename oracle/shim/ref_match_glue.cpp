// C-ABI glue around LINE RANGES of the reference's own src/ORBmatcher.cc and src/Frame.cc, compiled unmodified
// against oracle/shim/match_shim.h → oracle/_ref/libref_match.so.  TEST INFRASTRUCTURE, NOT PRODUCT.
//
// The *.inc files are cut out of /root/reference by `sed -n 'A,Bp'` at build time (oracle/Makefile), included below
// and deleted again; they never enter the repository.  Ranges:
//   src/ORBmatcher.cc:35-41      TH_HIGH / TH_LOW / HISTO_LENGTH, constructor
//   src/ORBmatcher.cc:43-221     SearchByProjection(Frame&, vector<MapPoint*>&, th, bFarPoints, thFarPoints) — the
//                                level-aware best/second-best loop of :84-140 — and RadiusByViewingCos
//   src/ORBmatcher.cc:222-425    SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) — the candidate-list top-2 loop of :264-325 inside its ordered walk
//   src/ORBmatcher.cc:644-759    SearchForInitialization
//   src/ORBmatcher.cc:760-901    SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)
//   src/ORBmatcher.cc:2008-2070  ComputeThreeMaxima, DescriptorDistance
//   src/Frame.cc:387-418         AssignFeaturesToGrid
//   src/Frame.cc:659-738         GetFeaturesInArea, PosInGrid
//   src/Frame.cc:862-914         the association tail of ComputeStereoMatches (body fragment, wrapped below)
#include "match_shim.h"

namespace ORB_SLAM3 {
#include "gen/orbmatcher_35_41.inc"
#include "gen/orbmatcher_43_221.inc"
#include "gen/orbmatcher_222_425.inc"
#include "gen/orbmatcher_644_759.inc"
#include "gen/orbmatcher_760_901.inc"
#include "gen/orbmatcher_2008_2070.inc"
#include "gen/frame_387_418.inc"
#include "gen/frame_659_738.inc"

// Frame::ComputeStereoMatches (src/Frame.cc:813-915) with the LightGlue call of :822-860 replaced by its result
// (`matches`); :816-817 restated, :862-914 is the reference's text.
void Frame::RefStereoTail(const std::vector<cv::DMatch> &matches) {
    mvuRight = vector<float>(N, -1.0f);
    mvDepth = vector<float>(N, -1.0f);
#include "gen/frame_862_914.inc"
}
}  // namespace ORB_SLAM3

using ORB_SLAM3::Frame;
using ORB_SLAM3::MapPoint;
using ORB_SLAM3::ORBmatcher;

static cv::Mat wrap_desc(const uint8_t *d, int n) { return cv::Mat(n, 32, CV_8UC1, (void *)d, 32); }

// Frame members as the constructors set them (src/Frame.cc:253-254, :344-345) + AssignFeaturesToGrid (:387-418)
static void fill_frame(Frame &F, const orc_keypoint *kps, const uint8_t *desc, int n, const float *bounds /*minX,minY,maxX,maxY*/) {
    F.N = n;
    F.Nleft = -1;
    F.mvKeysUn.resize(n);
    for (int i = 0; i < n; ++i) memcpy(&F.mvKeysUn[i], &kps[i], sizeof(cv::KeyPoint));
    F.mvKeys = F.mvKeysUn;
    if (desc) F.mDescriptors = wrap_desc(desc, n);
    F.mvpMapPoints.assign(n, nullptr);
    F.mvuRight.assign(n, -1.0f);
    F.mnMinX = bounds[0]; F.mnMinY = bounds[1]; F.mnMaxX = bounds[2]; F.mnMaxY = bounds[3];
    F.mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(F.mnMaxX - F.mnMinX);
    F.mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(F.mnMaxY - F.mnMinY);
    F.AssignFeaturesToGrid();
}

extern "C" {

int refm_constants(int32_t *th_low, int32_t *th_high, int32_t *histo_length) {
    *th_low = ORBmatcher::TH_LOW; *th_high = ORBmatcher::TH_HIGH; *histo_length = ORBmatcher::HISTO_LENGTH;
    return 0;
}

int refm_descriptor_distance(const uint8_t *a, const uint8_t *b) {
    return ORBmatcher::DescriptorDistance(wrap_desc(a, 1), wrap_desc(b, 1));
}

// counts[L] = histogram list sizes; ind[3] = ind1, ind2, ind3 (callers initialise them to -1 like :732-734)
void refm_three_maxima(const int32_t *counts, int L, int32_t *ind) {
    std::vector<std::vector<int>> h(L);
    for (int i = 0; i < L; ++i) h[i].assign(counts[i], 0);
    int a = -1, b = -1, c = -1;
    ORBmatcher m;
    m.ComputeThreeMaxima(h.data(), L, a, b, c);
    ind[0] = a; ind[1] = b; ind[2] = c;
}

int refm_features_in_area(const orc_keypoint *kps, int n, const float *bounds, const float *queries_xyr, int nq, int min_level, int max_level,
                          int32_t *cand_off, int32_t *cand, int cap) {
    Frame F;
    fill_frame(F, kps, nullptr, n, bounds);
    int total = 0;
    cand_off[0] = 0;
    for (int q = 0; q < nq; ++q) {
        const vector<size_t> v = F.GetFeaturesInArea(queries_xyr[3 * q], queries_xyr[3 * q + 1], queries_xyr[3 * q + 2], min_level, max_level);
        for (size_t j : v) { if (total < cap) cand[total] = (int32_t)j; ++total; }
        cand_off[q + 1] = total;
    }
    return total;
}

// prev_xy: vbPrevMatched (n1×2, updated in place like :754-756); matches12[n1] = vnMatches12; returns nmatches
int refm_search_init(const orc_keypoint *kps1, const uint8_t *desc1, int n1, const orc_keypoint *kps2, const uint8_t *desc2, int n2, const float *bounds,
                     float *prev_xy, int window, float nnratio, int check_ori, int32_t *matches12) {
    Frame F1, F2;
    fill_frame(F1, kps1, desc1, n1, bounds);
    fill_frame(F2, kps2, desc2, n2, bounds);
    std::vector<cv::Point2f> prev(n1);
    for (int i = 0; i < n1; ++i) prev[i] = cv::Point2f(prev_xy[2 * i], prev_xy[2 * i + 1]);
    std::vector<int> m12;
    ORBmatcher m(nnratio, check_ori != 0);
    const int nm = m.SearchForInitialization(F1, F2, prev, m12, window);
    for (int i = 0; i < n1; ++i) { matches12[i] = m12[i]; prev_xy[2 * i] = prev[i].x; prev_xy[2 * i + 1] = prev[i].y; }
    return nm;
}

// SearchByProjection(Frame&, vector<MapPoint*>&, th, bFarPoints, thFarPoints), monocular / rectified-stereo frames (Nleft == -1).
// Frame side: keypoints (mvKeysUn), descriptors, u_right (mvuRight, may be NULL = all -1), kp_obs[n] = Observations() of the map point
// already attached to keypoint i, or -1 for a null pointer.  Map-point side (m rows): proj = {mTrackProjX, mTrackProjY, mTrackProjXR,
// mTrackViewCos, mTrackDepth}, level = mnTrackScaleLevel, flags bit0 = mbTrackInView, bit1 = isBad(), n_obs = Observations(), 32-byte
// descriptors.  assigned[n] = index of the map point written into mvpMapPoints[i] by the call, or -1.  Returns nmatches.
int refm_search_by_projection(const orc_keypoint *kps, const uint8_t *desc, int n, const float *u_right, const int32_t *kp_obs, const float *bounds,
                              const float *scale_factors, int n_levels, const float *mp_proj5, const int32_t *mp_level, const uint8_t *mp_flags,
                              const int32_t *mp_obs, const uint8_t *mp_desc, int m, float nnratio, float th, int far_points, float th_far,
                              int32_t *assigned) {
    Frame F;
    fill_frame(F, kps, desc, n, bounds);
    if (u_right) F.mvuRight.assign(u_right, u_right + n);
    F.mvScaleFactors.assign(scale_factors, scale_factors + n_levels);
    std::vector<MapPoint> old(n);
    for (int i = 0; i < n; ++i)
        if (kp_obs[i] >= 0) { old[i].nObs = kp_obs[i]; F.mvpMapPoints[i] = &old[i]; }
    std::vector<MapPoint> mps(m);
    std::vector<MapPoint *> ptrs(m);
    for (int j = 0; j < m; ++j) {
        MapPoint &p = mps[j];
        p.mTrackProjX = mp_proj5[5 * j]; p.mTrackProjY = mp_proj5[5 * j + 1]; p.mTrackProjXR = mp_proj5[5 * j + 2];
        p.mTrackViewCos = mp_proj5[5 * j + 3]; p.mTrackDepth = mp_proj5[5 * j + 4];
        p.mnTrackScaleLevel = mp_level[j];
        p.mbTrackInView = mp_flags[j] & 1; p.bad = (mp_flags[j] & 2) != 0;
        p.nObs = mp_obs[j];
        p.desc = wrap_desc(mp_desc + (size_t)j * 32, 1);
        ptrs[j] = &p;
    }
    ORBmatcher matcher(nnratio, true);
    const int nm = matcher.SearchByProjection(F, ptrs, th, far_points != 0, th_far);
    for (int i = 0; i < n; ++i) {
        MapPoint *p = F.mvpMapPoints[i];
        assigned[i] = (p && p >= mps.data() && p < mps.data() + m) ? (int32_t)(p - mps.data()) : -1;
    }
    return nm;
}

// SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&), frames with Nleft == -1 and one camera.  Both feature vectors arrive as CSR
// (ascending node ids, per node the feature indices in their stored order).  kf_mp[i]: 0 = no map point at keyframe feature i,
// 1 = a good one, 2 = a bad one (isBad()).  assigned[n_f] = keyframe feature whose map point ended in vpMapPointMatches[i], or -1.
static void fill_fv(DBoW2::FeatureVector &fv, const int32_t *nodes, const int32_t *off, const int32_t *idx, int nn) {
    for (int k = 0; k < nn; ++k) fv[(unsigned)nodes[k]] = std::vector<unsigned int>(idx + off[k], idx + off[k + 1]);
}
int refm_search_by_bow(const orc_keypoint *kf_kps, const uint8_t *kf_desc, int n_kf, const uint8_t *kf_mp, const int32_t *kf_nodes,
                       const int32_t *kf_off, const int32_t *kf_idx, int kf_nn, const orc_keypoint *f_kps, const uint8_t *f_desc, int n_f,
                       const int32_t *f_nodes, const int32_t *f_off, const int32_t *f_idx, int f_nn, float nnratio, int check_ori,
                       int32_t *assigned) {
    const float bounds[4] = {0.f, 0.f, 640.f, 480.f};      // the grid plays no part here
    Frame F;
    fill_frame(F, f_kps, f_desc, n_f, bounds);
    fill_fv(F.mFeatVec, f_nodes, f_off, f_idx, f_nn);
    ORB_SLAM3::KeyFrame KF;
    KF.mvKeysUn.resize(n_kf);
    for (int i = 0; i < n_kf; ++i) memcpy(&KF.mvKeysUn[i], &kf_kps[i], sizeof(cv::KeyPoint));
    KF.mvKeys = KF.mvKeysUn;
    KF.mDescriptors = wrap_desc(kf_desc, n_kf);
    fill_fv(KF.mFeatVec, kf_nodes, kf_off, kf_idx, kf_nn);
    std::vector<MapPoint> mps(n_kf);
    KF.mps.assign(n_kf, nullptr);
    for (int i = 0; i < n_kf; ++i)
        if (kf_mp[i]) { mps[i].bad = kf_mp[i] == 2; KF.mps[i] = &mps[i]; }
    std::vector<MapPoint *> matches;
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int nm = matcher.SearchByBoW(&KF, F, matches);
    for (int i = 0; i < n_f; ++i) assigned[i] = matches[i] ? (int32_t)(matches[i] - mps.data()) : -1;
    return nm;
}

// SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12) (one camera per keyframe).  mp1 / mp2 as kf_mp above;
// matches12[n1] = feature of keyframe 2 whose map point ended in vpMatches12[i], or -1.
static void fill_kf(ORB_SLAM3::KeyFrame &KF, std::vector<MapPoint> &mps, const orc_keypoint *kps, const uint8_t *desc, int n, const uint8_t *mp,
                    const int32_t *nodes, const int32_t *off, const int32_t *idx, int nn) {
    KF.mvKeysUn.resize(n);
    for (int i = 0; i < n; ++i) memcpy(&KF.mvKeysUn[i], &kps[i], sizeof(cv::KeyPoint));
    KF.mvKeys = KF.mvKeysUn;
    KF.mDescriptors = wrap_desc(desc, n);
    fill_fv(KF.mFeatVec, nodes, off, idx, nn);
    mps.assign(n, MapPoint());
    KF.mps.assign(n, nullptr);
    for (int i = 0; i < n; ++i)
        if (mp[i]) { mps[i].bad = mp[i] == 2; KF.mps[i] = &mps[i]; }
}
int refm_search_by_bow_kf(const orc_keypoint *kps1, const uint8_t *desc1, int n1, const uint8_t *mp1, const int32_t *nodes1, const int32_t *off1,
                          const int32_t *idx1, int nn1, const orc_keypoint *kps2, const uint8_t *desc2, int n2, const uint8_t *mp2,
                          const int32_t *nodes2, const int32_t *off2, const int32_t *idx2, int nn2, float nnratio, int check_ori, int32_t *matches12) {
    ORB_SLAM3::KeyFrame K1, K2;
    std::vector<MapPoint> m1, m2;
    fill_kf(K1, m1, kps1, desc1, n1, mp1, nodes1, off1, idx1, nn1);
    fill_kf(K2, m2, kps2, desc2, n2, mp2, nodes2, off2, idx2, nn2);
    std::vector<MapPoint *> matches;
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int nm = matcher.SearchByBoW(&K1, &K2, matches);
    for (int i = 0; i < n1; ++i) matches12[i] = matches[i] ? (int32_t)(matches[i] - m2.data()) : -1;
    return nm;
}

// Tail of ComputeStereoMatches: matches = n_matches × (iL, iR, match.distance); u_left / u_right = keypoint pt.x of both images.
int refm_stereo_tail(const float *u_left, int n_left, const float *u_right, int n_right, const int32_t *iL, const int32_t *iR, const float *match_distance,
                     int n_matches, float mbf, float mb, float *mvu_right, float *mv_depth) {
    Frame F;
    F.N = n_left;
    F.mvKeys.resize(n_left); F.mvKeysRight.resize(n_right);
    for (int i = 0; i < n_left; ++i) F.mvKeys[i].pt.x = u_left[i];
    for (int i = 0; i < n_right; ++i) F.mvKeysRight[i].pt.x = u_right[i];
    F.mbf = mbf; F.mb = mb;
    std::vector<cv::DMatch> matches(n_matches);
    for (int i = 0; i < n_matches; ++i) matches[i] = cv::DMatch(iL[i], iR[i], match_distance[i]);
    F.RefStereoTail(matches);
    int kept = 0;
    for (int i = 0; i < n_left; ++i) { mvu_right[i] = F.mvuRight[i]; mv_depth[i] = F.mvDepth[i]; kept += F.mvuRight[i] != -1.0f; }
    return kept;
}

}  // extern "C"
