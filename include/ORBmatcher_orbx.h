/* ORBmatcher_orbx.h — the Hamming inner loops of the reference's ORBmatcher (src/ORBmatcher.cc) and of
 * Frame::ComputeStereoFishEyeMatches (src/Frame.cc:1060-1100) on the GPU, as a small C++ class over orbx.h.
 *
 * The reference-named class ORB_SLAM3::ORBmatcher (constructor, DescriptorDistance(cv::Mat, cv::Mat), SearchForInitialization,
 * SearchByProjection over Frame / MapPoint) is include/ORBmatcher.h; this header holds the array-level building blocks.
 *
 * Scope (SURVEY.md §8a rows a11–a15): DescriptorDistance, best/second-best search over candidate lists,
 * ratio tests, the rotation-histogram filter and the brute-force k=2 kNN.  The eleven SearchBy… and Fuse entry
 * points keep their projection geometry in the reference's src/ORBmatcher.cc (out of scope: pointer-chasing
 * over Frame/KeyFrame/MapPoint); INTEGRATION.md shows how their inner loops call into this class.
 * Matchers are stack objects in the reference (src/Tracking.cc:2511 …), so the device context is a
 * thread-local singleton: constructing an ORBmatcherDevice is free after the first use on a thread.
 */
#ifndef ORBMATCHER_ORBX_H
#define ORBMATCHER_ORBX_H

#include <climits>
#include <cstdlib>
#include <vector>

#include "orbx.h"

namespace ORB_SLAM3
{

class ORBmatcherDevice
{
public:
    static const int TH_HIGH = 100;      // src/ORBmatcher.cc:35
    static const int TH_LOW = 50;        // :36
    static const int HISTO_LENGTH = 30;  // :37

    explicit ORBmatcherDevice(float nnratio=0.6, bool checkOri=true): mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

    // ORBmatcher::DescriptorDistance (:2054-2070) on two 32-byte descriptors.
    static int DescriptorDistance(const unsigned char* a, const unsigned char* b) { return orbx_descriptor_distance(a, b); }

    // BFMatcher(NORM_HAMMING).knnMatch(query, train, k=2) (src/Frame.cc:1078): idx/dist sized 2*nq.
    bool KnnMatch2(const unsigned char* query, int nq, const unsigned char* train, long long ntrain,
                   std::vector<int>& idx, std::vector<int>& dist)
    {
        idx.assign(2*(size_t)nq, -1); dist.assign(2*(size_t)nq, INT_MAX);
        return orbx_hamming_knn2(Ctx(), query, nq, train, ntrain, idx.data(), dist.data()) == ORBX_OK;
    }

    // Lowe ratio of src/Frame.cc:1085 on the kNN result.
    bool RatioTest(const std::vector<int>& dist, double ratio, std::vector<unsigned char>& keep)
    {
        keep.assign(dist.size()/2, 0);
        return orbx_ratio_test(Ctx(), dist.data(), (int)(dist.size()/2), ratio, keep.data()) == ORBX_OK;
    }

    // best / second-best distances over per-query candidate lists (the loops at :84-140, :273-325, :812-864).
    bool Top2OverCandidates(const unsigned char* query, int nq, const unsigned char* train, long long ntrain,
                            const std::vector<int>& cand, const std::vector<int>& candOffsets,
                            std::vector<int>& bestIdx, std::vector<int>& bestDist, std::vector<int>& secondIdx, std::vector<int>& secondDist)
    {
        // secondIdx → the keypoint whose octave is bestLevel2 in the level-aware ratio rule (:101-128)
        bestIdx.assign(nq, -1); bestDist.assign(nq, 256); secondIdx.assign(nq, -1); secondDist.assign(nq, 256);
        return orbx_hamming_top2_lists(Ctx(), query, nq, train, ntrain, cand.data(), candOffsets.data(),
                                       bestIdx.data(), bestDist.data(), secondIdx.data(), secondDist.data()) == ORBX_OK;
    }

    // Rotation-consistency filter (:345-352 + ComputeThreeMaxima :2008-2049): keep[i]=0 for matches to drop.
    bool RotationFilter(const std::vector<float>& anglesA, const std::vector<float>& anglesB, std::vector<unsigned char>& keep)
    {
        keep.assign(anglesA.size(), 1);
        if(!mbCheckOrientation) return true;
        return orbx_rot_hist_filter(Ctx(), anglesA.data(), anglesB.data(), (int)anglesA.size(), keep.data()) == ORBX_OK;
    }

    float mfNNratio;
    bool mbCheckOrientation;

private:
    struct Holder {
        orbx_matcher* m;
        Holder(){ const char* d = std::getenv("ORBX_DEVICE"); m = orbx_matcher_create(d ? std::atoi(d) : 0); }
        ~Holder(){ if(m) orbx_matcher_destroy(m); }
    };
    static orbx_matcher* Ctx(){ static thread_local Holder h; return h.m; }
};

} // namespace ORB_SLAM3

#endif
