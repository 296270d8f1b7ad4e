/* ORBmatcher.h — drop-in replacement for /root/reference/include/ORBmatcher.h.
 *
 * Same namespace, class name, constructor, constants and member signatures as the reference
 * (include/ORBmatcher.h:36-103), so Tracking, LocalMapping and LoopClosing compile unchanged.  The members on
 * the hot path are defined HERE, inline, and forward to the sm_100a CUDA library
 * through the C ABI of orbx.h:
 *
 *   ORBmatcher(nnratio, checkOri)                                   src/ORBmatcher.cc:39-41
 *   static DescriptorDistance(const cv::Mat&, const cv::Mat&)       src/ORBmatcher.cc:2054-2070   (host popcount, no launch)
 *   SearchByProjection(Frame&, const vector<MapPoint*>&, th, …)     src/ORBmatcher.cc:43-213      → orbx_search_by_projection
 *   SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)              src/ORBmatcher.cc:222-425     → orbx_search_by_bow
 *   SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)           src/ORBmatcher.cc:760-901     → orbx_search_by_bow_keyframes
 *   SearchForInitialization(Frame&, Frame&, …, windowSize)          src/ORBmatcher.cc:644-759     → orbx_search_for_initialization_frames
 *   TH_LOW / TH_HIGH / HISTO_LENGTH                                 src/ORBmatcher.cc:35-37
 *
 * The other members (SearchForTriangulation, SearchBySim3, Fuse and the three remaining
 * SearchByProjection overloads) are projection geometry over KeyFrame / MapPoint pointers and stay in the
 * reference's own src/ORBmatcher.cc (out of scope); they are only DECLARED here, exactly as in the
 * reference, and their inner loops call the DescriptorDistance above.  INTEGRATION.md shows the five ranges of
 * src/ORBmatcher.cc a maintainer fences off (`#ifndef ORBX_DROPIN`) so that each function has one definition.
 *
 * Matchers are stack objects constructed per call in the reference (src/Tracking.cc:2511,2747 …), so the
 * constructor does nothing; the device context is a thread-local singleton created on first use (ORBX_DEVICE
 * selects the CUDA device).  There is no CPU fallback: a missing device or a CUDA error throws std::runtime_error
 * (the reference has no recoverable path either).
 *
 * SearchByProjection covers frames with Nleft == -1 (monocular, rectified stereo, RGB-D — every configuration
 * BASELINE.json names).  For the two-camera fisheye rig (Nleft != -1, src/ORBmatcher.cc:144-210) define
 * ORBX_KEEP_REFERENCE_TWO_CAMERA_PATH and keep the reference's body under the name SearchByProjectionTwoCameras;
 * without it such a frame throws instead of silently computing something else.
 */
#ifndef ORBMATCHER_H
#define ORBMATCHER_H

#include <cstdlib>
#include <cstring>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include <opencv2/core/core.hpp>

#include "orbx.h"

#ifndef ORBX_MATCHER_HOT_PATH_ONLY
#include <opencv2/features2d/features2d.hpp>
#include "sophus/sim3.hpp"

#include "MapPoint.h"
#include "KeyFrame.h"
#endif
#include "Frame.h"

namespace ORB_SLAM3
{

class ORBmatcher
{
public:

    ORBmatcher(float nnratio=0.6, bool checkOri=true): mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

    // Computes the Hamming distance between two ORB descriptors
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b)
    {
        return orbx_descriptor_distance(a.ptr<unsigned char>(), b.ptr<unsigned char>());
    }

    // Search matches between Frame keypoints and projected MapPoints. Returns number of matches
    // Used to track the local map (Tracking)
    int SearchByProjection(Frame &F, const std::vector<MapPoint*> &vpMapPoints, const float th=3, const bool bFarPoints = false, const float thFarPoints = 50.0f)
    {
        if(F.Nleft != -1)
        {
#ifdef ORBX_KEEP_REFERENCE_TWO_CAMERA_PATH
            return SearchByProjectionTwoCameras(F, vpMapPoints, th, bFarPoints, thFarPoints);
#else
            throw std::runtime_error("ORBmatcher::SearchByProjection: two-camera frames (Nleft != -1) are not on the device path");
#endif
        }
        const int n = (int)F.mvKeysUn.size(), m = (int)vpMapPoints.size();
        if(n == 0 || m == 0) return 0;
        std::vector<int> kpObs(n, -1), level(m), obs(m), assigned(n, -1);
        for(int i=0; i<n; i++)
            if(F.mvpMapPoints[i]) kpObs[i] = F.mvpMapPoints[i]->Observations();
        std::vector<float> proj(5*(size_t)m);
        std::vector<unsigned char> flags(m), desc(32*(size_t)m);
        for(int j=0; j<m; j++)
        {
            MapPoint* pMP = vpMapPoints[j];
            const bool inView = pMP->mbTrackInView;              // :52 (mbTrackInViewR only matters when Nleft != -1)
            const bool bad = inView && pMP->isBad();             // :58, evaluated like the reference: only for points in view
            flags[j] = (unsigned char)((inView ? 1 : 0) | (bad ? 2 : 0));
            proj[5*j] = pMP->mTrackProjX; proj[5*j+1] = pMP->mTrackProjY; proj[5*j+2] = pMP->mTrackProjXR;
            proj[5*j+3] = pMP->mTrackViewCos; proj[5*j+4] = pMP->mTrackDepth;
            level[j] = pMP->mnTrackScaleLevel;
            obs[j] = inView ? pMP->Observations() : 0;
            if(inView && !bad)
            {
                const cv::Mat d = pMP->GetDescriptor();
                std::memcpy(&desc[32*(size_t)j], d.ptr<unsigned char>(), 32);
            }
        }
        const float bounds[4] = {F.mnMinX, F.mnMinY, F.mnMaxX, F.mnMaxY};
        std::vector<unsigned char> rows;
        int nmatches = 0;
        Check(orbx_search_by_projection(Ctx(), reinterpret_cast<const orbx_keypoint*>(F.mvKeysUn.data()), Rows(F.mDescriptors, n, rows), n,
                                        F.mvuRight.empty() ? nullptr : F.mvuRight.data(), kpObs.data(), bounds, F.mvScaleFactors.data(),
                                        (int)F.mvScaleFactors.size(), proj.data(), level.data(), flags.data(), obs.data(), desc.data(), m,
                                        mfNNratio, th, bFarPoints ? 1 : 0, thFarPoints, assigned.data(), &nmatches),
              "SearchByProjection");
        for(int i=0; i<n; i++)
            if(assigned[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[assigned[i]];   // :130
        return nmatches;
    }

    // Matching for the Map Initialization (only used in the monocular case)
    int SearchForInitialization(Frame &F1, Frame &F2, std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize=10)
    {
        const int n1 = (int)F1.mvKeysUn.size(), n2 = (int)F2.mvKeysUn.size();
        vnMatches12 = std::vector<int>(n1, -1);
        if(n1 == 0 || n2 == 0) return 0;
        const float bounds[4] = {F2.mnMinX, F2.mnMinY, F2.mnMaxX, F2.mnMaxY};
        std::vector<unsigned char> rows1, rows2;
        int nmatches = 0;
        static_assert(sizeof(cv::Point2f) == 8 && sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "layout");
        Check(orbx_search_for_initialization_frames(Ctx(), reinterpret_cast<const orbx_keypoint*>(F1.mvKeysUn.data()), Rows(F1.mDescriptors, n1, rows1), n1,
                                                    reinterpret_cast<const orbx_keypoint*>(F2.mvKeysUn.data()), Rows(F2.mDescriptors, n2, rows2), n2, bounds,
                                                    reinterpret_cast<float*>(vbPrevMatched.data()), windowSize, mfNNratio, mbCheckOrientation ? 1 : 0,
                                                    vnMatches12.data(), &nmatches),
              "SearchForInitialization");
        return nmatches;
    }

#if !defined(ORBX_MATCHER_HOT_PATH_ONLY) || defined(ORBX_SHIM_FRAME_H)   // needs the complete KeyFrame type (KeyFrame.h, or the test stand-in)
    // Search matches between MapPoints in a KeyFrame and ORB in a Frame.
    // Brute force constrained to ORB that belong to the same vocabulary node (at a certain level)
    // Used in Relocalisation and Loop Detection
    int SearchByBoW(KeyFrame *pKF, Frame &F, std::vector<MapPoint*> &vpMapPointMatches)
    {
        if(F.Nleft != -1 || pKF->mpCamera2)
            throw std::runtime_error("ORBmatcher::SearchByBoW: two-camera frames (Nleft != -1) are not on the device path");
        const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
        const int nKF = (int)vpMapPointsKF.size(), nF = F.N;
        vpMapPointMatches = std::vector<MapPoint*>(nF, static_cast<MapPoint*>(NULL));      // :226
        if(nKF == 0 || nF == 0) return 0;
        std::vector<unsigned char> mp(nKF);
        std::vector<float> angKF(nKF), angF(nF);
        for(int i=0; i<nKF; i++)
        {
            MapPoint* pMP = vpMapPointsKF[i];
            mp[i] = !pMP ? 0 : (pMP->isBad() ? 2 : 1);                                     // :256-260
            angKF[i] = pKF->mvKeysUn[i].angle;                                              // :338
        }
        for(int i=0; i<nF; i++) angF[i] = F.mvKeys[i].angle;                                // :345
        std::vector<int> kn, ko(1, 0), ki, fn, fo(1, 0), fi;
        Flatten(pKF->mFeatVec, kn, ko, ki);
        Flatten(F.mFeatVec, fn, fo, fi);
        std::vector<unsigned char> rowsKF, rowsF;
        std::vector<int> assigned(nF, -1);
        int nmatches = 0;
        Check(orbx_search_by_bow(Ctx(), Rows(pKF->mDescriptors, nKF, rowsKF), angKF.data(), nKF, mp.data(), kn.data(), ko.data(), ki.data(), (int)kn.size(),
                                 Rows(F.mDescriptors, nF, rowsF), angF.data(), nF, fn.data(), fo.data(), fi.data(), (int)fn.size(), mfNNratio,
                                 mbCheckOrientation ? 1 : 0, assigned.data(), &nmatches),
              "SearchByBoW");
        for(int i=0; i<nF; i++)
            if(assigned[i] >= 0) vpMapPointMatches[i] = vpMapPointsKF[assigned[i]];        // :335
        return nmatches;
    }

    // Matching for triangulating / merging between two keyframes constrained to the same vocabulary node (Loop Closing, Merging)
    int SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<MapPoint*> &vpMatches12)
    {
        if(pKF1->NLeft != -1 || pKF2->NLeft != -1)
            throw std::runtime_error("ORBmatcher::SearchByBoW: two-camera keyframes (NLeft != -1) are not on the device path");
        const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
        const int n1 = (int)vpMapPoints1.size(), n2 = (int)vpMapPoints2.size();
        vpMatches12 = std::vector<MapPoint*>(n1, static_cast<MapPoint*>(NULL));            // :772
        if(n1 == 0 || n2 == 0) return 0;
        std::vector<unsigned char> mp1(n1), mp2(n2);
        std::vector<float> ang1(n1), ang2(n2);
        for(int i=0; i<n1; i++){ MapPoint* p = vpMapPoints1[i]; mp1[i] = !p ? 0 : (p->isBad() ? 2 : 1); ang1[i] = pKF1->mvKeysUn[i].angle; }
        for(int i=0; i<n2; i++){ MapPoint* p = vpMapPoints2[i]; mp2[i] = !p ? 0 : (p->isBad() ? 2 : 1); ang2[i] = pKF2->mvKeysUn[i].angle; }
        std::vector<int> an, ao(1, 0), ai, bn, bo(1, 0), bi;
        Flatten(pKF1->mFeatVec, an, ao, ai);
        Flatten(pKF2->mFeatVec, bn, bo, bi);
        std::vector<unsigned char> rows1, rows2;
        std::vector<int> m12(n1, -1);
        int nmatches = 0;
        Check(orbx_search_by_bow_keyframes(Ctx(), Rows(pKF1->mDescriptors, n1, rows1), ang1.data(), n1, mp1.data(), an.data(), ao.data(), ai.data(), (int)an.size(),
                                           Rows(pKF2->mDescriptors, n2, rows2), ang2.data(), n2, mp2.data(), bn.data(), bo.data(), bi.data(), (int)bn.size(),
                                           mfNNratio, mbCheckOrientation ? 1 : 0, m12.data(), &nmatches),
              "SearchByBoW");
        for(int i=0; i<n1; i++)
            if(m12[i] >= 0) vpMatches12[i] = vpMapPoints2[m12[i]];                         // :847
        return nmatches;
    }
#endif

#ifndef ORBX_MATCHER_HOT_PATH_ONLY
    // ---- out of scope: declared exactly as in the reference, defined by the reference's src/ORBmatcher.cc ----
    int SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono);
    int SearchByProjection(Frame &CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*> &sAlreadyFound, const float th, const int ORBdist);
    int SearchByProjection(KeyFrame* pKF, Sophus::Sim3<float> &Scw, const std::vector<MapPoint*> &vpPoints, std::vector<MapPoint*> &vpMatched, int th, float ratioHamming=1.0);
    int SearchByProjection(KeyFrame* pKF, Sophus::Sim3<float> &Scw, const std::vector<MapPoint*> &vpPoints, const std::vector<KeyFrame*> &vpPointsKFs, std::vector<MapPoint*> &vpMatched, std::vector<KeyFrame*> &vpMatchedKF, int th, float ratioHamming=1.0);
    int SearchForTriangulation(KeyFrame *pKF1, KeyFrame* pKF2,
                               std::vector<std::pair<size_t, size_t> > &vMatchedPairs, const bool bOnlyStereo, const bool bCoarse = false);
    int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint *> &vpMatches12, const Sophus::Sim3f &S12, const float th);
    int Fuse(KeyFrame* pKF, const std::vector<MapPoint *> &vpMapPoints, const float th=3.0, const bool bRight = false);
    int Fuse(KeyFrame* pKF, Sophus::Sim3f &Scw, const std::vector<MapPoint*> &vpPoints, float th, std::vector<MapPoint *> &vpReplacePoint);
#endif

public:

    static const int TH_LOW = 50;         // src/ORBmatcher.cc:36
    static const int TH_HIGH = 100;       // src/ORBmatcher.cc:35
    static const int HISTO_LENGTH = 30;   // src/ORBmatcher.cc:37
#ifdef EIGEN_MAKE_ALIGNED_OPERATOR_NEW
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#endif

    // Last CUDA / argument error of the calling thread's device context ("" when none).
    static std::string LastError(){ orbx_matcher* c = CtxNoThrow(); return c ? orbx_matcher_last_error(c) : orbx_matcher_last_error(nullptr); }

protected:
#ifndef ORBX_MATCHER_HOT_PATH_ONLY
    float RadiusByViewingCos(const float &viewCos);
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int &ind1, int &ind2, int &ind3);
#endif
#ifdef ORBX_KEEP_REFERENCE_TWO_CAMERA_PATH
    int SearchByProjectionTwoCameras(Frame &F, const std::vector<MapPoint*> &vpMapPoints, const float th, const bool bFarPoints, const float thFarPoints);
#endif

    float mfNNratio;
    bool mbCheckOrientation;

private:
    struct Holder {
        orbx_matcher* m;
        Holder(){ const char* d = std::getenv("ORBX_DEVICE"); m = orbx_matcher_create(d ? std::atoi(d) : 0); }
        ~Holder(){ if(m) orbx_matcher_destroy(m); }
    };
    static orbx_matcher* CtxNoThrow(){ static thread_local Holder h; return h.m; }
    static orbx_matcher* Ctx()
    {
        orbx_matcher* c = CtxNoThrow();
        if(!c) throw std::runtime_error(std::string("ORBmatcher: ") + orbx_matcher_last_error(nullptr));
        return c;
    }
    static void Check(int rc, const char* what)
    {
        if(rc != ORBX_OK) throw std::runtime_error(std::string("ORBmatcher::") + what + ": " + orbx_matcher_last_error(CtxNoThrow()));
    }
    // DBoW2::FeatureVector (a std::map<NodeId, std::vector<unsigned int>>) as CSR: ascending node ids, offsets, feature indices
    template <class FV>
    static void Flatten(const FV &fv, std::vector<int> &nodes, std::vector<int> &off, std::vector<int> &idx)
    {
        for(typename FV::const_iterator it = fv.begin(); it != fv.end(); ++it)
        {
            nodes.push_back((int)it->first);
            idx.insert(idx.end(), it->second.begin(), it->second.end());
            off.push_back((int)idx.size());
        }
    }
    // n descriptor rows of 32 bytes as one contiguous block (mDescriptors is continuous in the reference; copy if a view is not)
    static const unsigned char* Rows(const cv::Mat &D, int n, std::vector<unsigned char> &scratch)
    {
        if(n == 0) return nullptr;
        if((size_t)D.step == 32) return D.ptr<unsigned char>();
        scratch.resize(32*(size_t)n);
        for(int i=0; i<n; i++) std::memcpy(&scratch[32*(size_t)i], D.ptr<unsigned char>(i), 32);
        return scratch.data();
    }
};

}// namespace ORB_SLAM

#endif // ORBMATCHER_H
