// Drives include/ORBextractor.h (the drop-in adapter, reference signatures) and include/ORBmatcher_orbx.h
// exactly like Frame::ExtractORB does (src/Frame.cc:420-427), against the opencv2 shim, and compares the
// result with the reference's own ORBextractor.cc (oracle/_ref) when that library is linked in.
// With a 4th argument (a vocabulary in the ORBvoc text format) it also runs include/ORBVocabulary_orbx.h like
// Frame::ComputeBoW (src/Frame.cc:739-747) and compares with the reference's own DBoW2 (oracle/_ref/libref_bow.so).
// It then fills two Frames from two extractions and calls ORB_SLAM3::ORBmatcher::SearchForInitialization / SearchByProjection of the
// drop-in include/ORBmatcher.h like src/Tracking.cc:2512 and :3447 do, comparing with the reference's own function bodies
// (oracle/_ref/libref_match.so, WITH_REF_MATCH).
// Usage: adapter_check <width> <height> <seed> [vocabulary.txt]   (prints "OK n mono" or a diagnostic; exit code 0/1)
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <map>
#include <vector>

#include "ORBextractor.h"
#include "ORBmatcher.h"        // the drop-in class (reference signatures), compiled with -DORBX_MATCHER_HOT_PATH_ONLY against tests/cpp/shim/Frame.h
#include "ORBmatcher_orbx.h"
#include "ORBVocabulary_orbx.h"

extern "C" {
void *ref_create(int, float, int, int, int);
void ref_destroy(void *);
int ref_extract(void *, const uint8_t *, int, int, size_t, const int32_t *, int, int, int, orc_keypoint *, uint8_t *, int, int *, int *);
#ifdef WITH_REF_MATCH
int refm_search_init(const orc_keypoint *, const uint8_t *, int, const orc_keypoint *, const uint8_t *, int, const float *, float *, int, float, int, int32_t *);
int refm_search_by_projection(const orc_keypoint *, const uint8_t *, int, const float *, const int32_t *, const float *, const float *, int, const float *,
                              const int32_t *, const uint8_t *, const int32_t *, const uint8_t *, int, float, float, int, float, int32_t *);
int refm_search_by_bow(const orc_keypoint *, const uint8_t *, int, const uint8_t *, const int32_t *, const int32_t *, const int32_t *, int, const orc_keypoint *,
                       const uint8_t *, int, const int32_t *, const int32_t *, const int32_t *, int, float, int, int32_t *);
int refm_search_by_bow_kf(const orc_keypoint *, const uint8_t *, int, const uint8_t *, const int32_t *, const int32_t *, const int32_t *, int, const orc_keypoint *,
                          const uint8_t *, int, const uint8_t *, const int32_t *, const int32_t *, const int32_t *, int, float, int, int32_t *);
#endif
#ifdef WITH_REF_BOW
void *ref_vocab_load_text(const char *);
void ref_vocab_free(void *);
int ref_bow_transform(void *, const uint8_t *, int, int, uint32_t *, double *, int *, uint32_t *, int32_t *, uint32_t *, int *);
#endif
}

int main(int argc, char **argv) {
    const int W = argc > 1 ? atoi(argv[1]) : 640, H = argc > 2 ? atoi(argv[2]) : 480;
    unsigned s = argc > 3 ? (unsigned)atoi(argv[3]) : 1u;
    cv::Mat img(H, W, CV_8UC1);
    // blocky random texture: plenty of corners
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            unsigned h = (unsigned)(x / 5) * 73856093u ^ (unsigned)(y / 5) * 19349663u ^ s * 83492791u;
            h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
            img.at<uchar>(y, x) = (uchar)(h & 0xff);
        }
    ORB_SLAM3::ORBextractor ex(1000, 1.2f, 8, 20, 7);
    ex.mvDynamicArea.push_back(cv::Rect2i(40, 30, 100, 80));
    std::vector<cv::KeyPoint> keys;
    cv::Mat desc, mask;
    std::vector<int> lap = {0, 1000};
    const int mono = ex(img, mask, keys, desc, lap);
    if (mono < 0) { printf("FAIL extractor returned %d\n", mono); return 1; }
    ex.MaterializePyramid();
    if (ex.mvImagePyramid[1].cols != cvRound((float)W * ex.GetInverseScaleFactors()[1])) { printf("FAIL pyramid size\n"); return 1; }

    void *ref = ref_create(1000, 1.2f, 8, 20, 7);
    std::vector<orc_keypoint> rk(1200);
    std::vector<uint8_t> rd(1200 * 32);
    int rn = 0, rmono = 0;
    const int32_t rect[4] = {40, 30, 100, 80};
    ref_extract(ref, img.data, H, W, (size_t)img.step, rect, 1, 0, 1000, rk.data(), rd.data(), 1200, &rn, &rmono);
    ref_destroy(ref);
    if (rn != (int)keys.size() || rmono != mono) { printf("FAIL count %d vs %zu, mono %d vs %d\n", rn, keys.size(), rmono, mono); return 1; }
    if (memcmp(rk.data(), keys.data(), (size_t)rn * 28) != 0) { printf("FAIL keypoints differ\n"); return 1; }
    for (int i = 0; i < rn; ++i)
        if (memcmp(&rd[(size_t)i * 32], desc.ptr(i), 32) != 0) { printf("FAIL descriptor %d differs\n", i); return 1; }

    // matcher: left/right style kNN on the extracted descriptors against themselves shifted by one
    ORB_SLAM3::ORBmatcherDevice m(0.7f, true);
    std::vector<int> idx, dist;
    if (!m.KnnMatch2(desc.ptr(0), rn, desc.ptr(0), rn, idx, dist)) { printf("FAIL knn\n"); return 1; }
    for (int i = 0; i < rn; ++i)
        if (dist[2 * i] != 0) { printf("FAIL self match %d\n", i); return 1; }
    if (ORB_SLAM3::ORBmatcherDevice::DescriptorDistance(desc.ptr(0), desc.ptr(0)) != 0) { printf("FAIL distance\n"); return 1; }
    // ---- the drop-in ORBmatcher on two Frames: frame 2 = the same scene moved by (7, 3) pixels ----
    {
        cv::Mat img2(H, W, CV_8UC1);
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) img2.at<uchar>(y, x) = img.at<uchar>(std::min(H - 1, std::max(0, y - 3)), std::min(W - 1, std::max(0, x - 7)));
        ORB_SLAM3::ORBextractor ex2(1000, 1.2f, 8, 20, 7);
        std::vector<cv::KeyPoint> keys2;
        cv::Mat desc2;
        std::vector<int> lap0 = {0, 0};
        if (ex2(img2, mask, keys2, desc2, lap0) < 0) { printf("FAIL second extraction\n"); return 1; }
        ORB_SLAM3::Frame F1, F2;
        ORB_SLAM3::Frame *Fs[2] = {&F1, &F2};
        for (int f = 0; f < 2; ++f) {
            ORB_SLAM3::Frame &F = *Fs[f];
            F.mvKeysUn = f == 0 ? keys : keys2; F.mvKeys = F.mvKeysUn;
            F.N = (int)F.mvKeysUn.size();
            F.mDescriptors = f == 0 ? desc : desc2;
            F.mvpMapPoints.assign(F.N, nullptr);
            F.mvuRight.assign(F.N, -1.0f);
            F.mnMinX = 0; F.mnMinY = 0; F.mnMaxX = (float)W; F.mnMaxY = (float)H;
            F.mvScaleFactors = ex.GetScaleFactors();
        }
        const float bounds[4] = {0, 0, (float)W, (float)H};
        (void)bounds;
        std::vector<cv::Point2f> prev(F1.N);
        for (int i = 0; i < F1.N; ++i) prev[i] = F1.mvKeysUn[i].pt;
        std::vector<cv::Point2f> prevRef = prev;
        std::vector<int> m12;
        ORB_SLAM3::ORBmatcher matcher(0.9f, true);                       // src/Tracking.cc:2511
        const int nm = matcher.SearchForInitialization(F1, F2, prev, m12, 100);
        if (nm < 20) { printf("FAIL SearchForInitialization found %d matches\n", nm); return 1; }
        if (ORB_SLAM3::ORBmatcher::DescriptorDistance(desc.row(0), desc.row(0)) != 0 || ORB_SLAM3::ORBmatcher::TH_LOW != 50) { printf("FAIL matcher statics\n"); return 1; }
#ifdef WITH_REF_MATCH
        std::vector<int32_t> rm12(F1.N);
        const int rnm = refm_search_init((const orc_keypoint *)F1.mvKeysUn.data(), desc.ptr(0), F1.N, (const orc_keypoint *)F2.mvKeysUn.data(), desc2.ptr(0), F2.N,
                                         bounds, (float *)prevRef.data(), 100, 0.9f, 1, rm12.data());
        if (rnm != nm || memcmp(rm12.data(), m12.data(), (size_t)F1.N * 4) != 0 || memcmp(prevRef.data(), prev.data(), (size_t)F1.N * 8) != 0) {
            printf("FAIL SearchForInitialization differs from the reference (%d vs %d)\n", nm, rnm); return 1;
        }
#endif
        // map points = frame-2 keypoints projected into frame 1 (two per keypoint region so that keypoints are contested)
        const int M = F2.N;
        std::vector<ORB_SLAM3::MapPoint> mps(M);
        std::vector<ORB_SLAM3::MapPoint *> vp(M);
        std::vector<float> p5(5 * (size_t)M);
        std::vector<int32_t> lvl(M), obs(M), kpObs(F1.N, -1);
        std::vector<uint8_t> flg(M);
        for (int j = 0; j < M; ++j) {
            ORB_SLAM3::MapPoint &p = mps[j];
            p.mTrackProjX = F2.mvKeysUn[j].pt.x - 7 + (j % 3) - 1; p.mTrackProjY = F2.mvKeysUn[j].pt.y - 3; p.mTrackProjXR = -1;
            p.mTrackViewCos = (j % 4) ? 0.9995f : 0.9f; p.mTrackDepth = 10.f + j % 50;
            p.mnTrackScaleLevel = F2.mvKeysUn[j].octave; p.mbTrackInView = (j % 17) != 0; p.bad = (j % 29) == 0; p.nObs = j % 3;
            p.desc = desc2.row(j);
            vp[j] = &p;
            p5[5 * j] = p.mTrackProjX; p5[5 * j + 1] = p.mTrackProjY; p5[5 * j + 2] = p.mTrackProjXR; p5[5 * j + 3] = p.mTrackViewCos; p5[5 * j + 4] = p.mTrackDepth;
            lvl[j] = p.mnTrackScaleLevel; obs[j] = p.nObs; flg[j] = (uint8_t)((p.mbTrackInView ? 1 : 0) | (p.bad ? 2 : 0));
        }
        ORB_SLAM3::ORBmatcher tracker(0.8f);                             // src/Tracking.cc:3447
        const int np = tracker.SearchByProjection(F1, vp, 3, true, 40.0f);
        if (np < 20) { printf("FAIL SearchByProjection found %d matches\n", np); return 1; }
#ifdef WITH_REF_MATCH
        std::vector<int32_t> asg(F1.N);
        const int rnp = refm_search_by_projection((const orc_keypoint *)F1.mvKeysUn.data(), desc.ptr(0), F1.N, nullptr, kpObs.data(), bounds, F1.mvScaleFactors.data(),
                                                  (int)F1.mvScaleFactors.size(), p5.data(), lvl.data(), flg.data(), obs.data(), desc2.ptr(0), M, 0.8f, 3.f, 1, 40.0f, asg.data());
        if (rnp != np) { printf("FAIL SearchByProjection count %d vs reference %d\n", np, rnp); return 1; }
        for (int i = 0; i < F1.N; ++i) {
            const int got = F1.mvpMapPoints[i] ? (int)(F1.mvpMapPoints[i] - mps.data()) : -1;
            if (got != asg[i]) { printf("FAIL SearchByProjection keypoint %d: map point %d vs reference %d\n", i, got, asg[i]); return 1; }
        }
#endif
    // SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (src/Tracking.cc:2747 / :3724 style): the first extraction as the keyframe, the second
    // as the frame, feature vectors = a coarse quantisation of the descriptors so that matching features mostly share a node
    {
        ORB_SLAM3::KeyFrame KF;
        ORB_SLAM3::Frame F;
        KF.mvKeysUn = keys; KF.mvKeys = keys; KF.mDescriptors = desc;
        F.mvKeysUn = keys2; F.mvKeys = keys2; F.N = (int)keys2.size(); F.mDescriptors = desc2;
        std::vector<ORB_SLAM3::MapPoint> mps(keys.size());
        KF.mps.assign(keys.size(), nullptr);
        std::vector<uint8_t> kfmp(keys.size());
        for (size_t i = 0; i < keys.size(); ++i) {
            kfmp[i] = (uint8_t)((i % 11) == 0 ? 0 : ((i % 23) == 0 ? 2 : 1));
            if (kfmp[i]) { mps[i].bad = kfmp[i] == 2; KF.mps[i] = &mps[i]; }
            KF.mFeatVec[(unsigned)(keys[i].octave * 8 + ((int)keys[i].pt.y * 8 / H))].push_back((unsigned)i);
        }
        for (size_t i = 0; i < keys2.size(); ++i) F.mFeatVec[(unsigned)(keys2[i].octave * 8 + ((int)keys2[i].pt.y * 8 / H))].push_back((unsigned)i);
        std::vector<ORB_SLAM3::MapPoint *> matches;
        ORB_SLAM3::ORBmatcher reloc(0.75f, true);                        // src/Tracking.cc:3685
        const int nb = reloc.SearchByBoW(&KF, F, matches);
        if (nb < 10 || (int)matches.size() != F.N) { printf("FAIL SearchByBoW found %d matches\n", nb); return 1; }
#ifdef WITH_REF_MATCH
        std::vector<int32_t> kn, ko(1, 0), ki, fn, fo(1, 0), fi, asg(F.N);
        for (auto &e : KF.mFeatVec) { kn.push_back((int32_t)e.first); ki.insert(ki.end(), e.second.begin(), e.second.end()); ko.push_back((int32_t)ki.size()); }
        for (auto &e : F.mFeatVec) { fn.push_back((int32_t)e.first); fi.insert(fi.end(), e.second.begin(), e.second.end()); fo.push_back((int32_t)fi.size()); }
        const int rnb = refm_search_by_bow((const orc_keypoint *)keys.data(), desc.ptr(0), (int)keys.size(), kfmp.data(), kn.data(), ko.data(), ki.data(), (int)kn.size(),
                                           (const orc_keypoint *)keys2.data(), desc2.ptr(0), F.N, fn.data(), fo.data(), fi.data(), (int)fn.size(), 0.75f, 1, asg.data());
        if (rnb != nb) { printf("FAIL SearchByBoW count %d vs reference %d\n", nb, rnb); return 1; }
        for (int i = 0; i < F.N; ++i) {
            const int got = matches[i] ? (int)(matches[i] - mps.data()) : -1;
            if (got != asg[i]) { printf("FAIL SearchByBoW frame feature %d: keyframe feature %d vs reference %d\n", i, got, asg[i]); return 1; }
        }
#endif
        // SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&) (src/LoopClosing.cc:592-593 style): the frame becomes a second keyframe
        ORB_SLAM3::KeyFrame KF2;
        KF2.mvKeysUn = keys2; KF2.mvKeys = keys2; KF2.mDescriptors = desc2; KF2.mFeatVec = F.mFeatVec;
        std::vector<ORB_SLAM3::MapPoint> mps2(keys2.size());
        KF2.mps.assign(keys2.size(), nullptr);
        std::vector<uint8_t> kfmp2(keys2.size());
        for (size_t i = 0; i < keys2.size(); ++i) {
            kfmp2[i] = (uint8_t)((i % 7) == 0 ? 0 : ((i % 19) == 0 ? 2 : 1));
            if (kfmp2[i]) { mps2[i].bad = kfmp2[i] == 2; KF2.mps[i] = &mps2[i]; }
        }
        std::vector<ORB_SLAM3::MapPoint *> m12;
        const int nk = reloc.SearchByBoW(&KF, &KF2, m12);
        if (nk < 10 || m12.size() != keys.size()) { printf("FAIL SearchByBoW(KF, KF) found %d matches\n", nk); return 1; }
#ifdef WITH_REF_MATCH
        std::vector<int32_t> r12(keys.size());
        const int rnk = refm_search_by_bow_kf((const orc_keypoint *)keys.data(), desc.ptr(0), (int)keys.size(), kfmp.data(), kn.data(), ko.data(), ki.data(), (int)kn.size(),
                                              (const orc_keypoint *)keys2.data(), desc2.ptr(0), (int)keys2.size(), kfmp2.data(), fn.data(), fo.data(), fi.data(),
                                              (int)fn.size(), 0.75f, 1, r12.data());
        if (rnk != nk) { printf("FAIL SearchByBoW(KF, KF) count %d vs reference %d\n", nk, rnk); return 1; }
        for (size_t i = 0; i < keys.size(); ++i) {
            const int got = m12[i] ? (int)(m12[i] - mps2.data()) : -1;
            if (got != r12[i]) { printf("FAIL SearchByBoW(KF, KF) feature %zu: %d vs reference %d\n", i, got, r12[i]); return 1; }
        }
#endif
    }
    }
    // vocabulary: BowVector / FeatureVector of the extracted descriptors, levelsup 4 as in Frame::ComputeBoW
    if (argc > 4) {
        typedef std::map<unsigned int, double> BowVector;                         // DBoW2::BowVector's base
        typedef std::map<unsigned int, std::vector<unsigned int> > FeatureVector; // DBoW2::FeatureVector's base
        ORB_SLAM3::ORBVocabularyDevice voc;
        if (!voc.loadFromTextFile(argv[4]) || voc.empty()) { printf("FAIL vocabulary load: %s\n", voc.LastError()); return 1; }
        BowVector bv; FeatureVector fv;
        if (!voc.transform(desc.ptr(0), rn, bv, fv, 4)) { printf("FAIL transform: %s\n", voc.LastError()); return 1; }
        if (bv.empty() || fv.empty()) { printf("FAIL empty bag of words\n"); return 1; }
#ifdef WITH_REF_BOW
        void *rv = ref_vocab_load_text(argv[4]);
        if (!rv) { printf("FAIL reference vocabulary load\n"); return 1; }
        std::vector<uint32_t> ids(rn), nodes(rn), fidx(rn);
        std::vector<double> vals(rn);
        std::vector<int32_t> off(rn + 1);
        int nb = 0, nf = 0;
        ref_bow_transform(rv, desc.ptr(0), rn, 4, ids.data(), vals.data(), &nb, nodes.data(), off.data(), fidx.data(), &nf);
        ref_vocab_free(rv);
        if (nb != (int)bv.size() || nf != (int)fv.size()) { printf("FAIL bow sizes %d/%zu %d/%zu\n", nb, bv.size(), nf, fv.size()); return 1; }
        int i = 0;
        for (BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++i)
            if (it->first != ids[i] || memcmp(&it->second, &vals[i], 8) != 0) { printf("FAIL bow entry %d\n", i); return 1; }
        i = 0;
        for (FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++i)
            if (it->first != nodes[i] || it->second != std::vector<unsigned int>(fidx.begin() + off[i], fidx.begin() + off[i + 1])) { printf("FAIL feature vector entry %d\n", i); return 1; }
#endif
    }
    printf("OK %d %d\n", rn, mono);
    return 0;
}
