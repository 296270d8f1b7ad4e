// Stand-in declarations so that LINE RANGES of the reference's own src/ORBmatcher.cc and src/Frame.cc compile here
// unmodified (oracle/Makefile target `_ref/libref_match.so`).  TEST INFRASTRUCTURE, NOT PRODUCT.
//
// The two files do not compile whole in this container (Eigen, Sophus, DBoW2, LibTorch, LightGlue),
// but the functions on the hot path only touch a handful of Frame / MapPoint members.  This header declares exactly
// those members, with the reference's names and types (include/Frame.h:52-53,118-120,258-260,295-298,337-352;
// include/MapPoint.h:123,132,146,171-179; include/ORBmatcher.h:36-103), and nothing else.  The bodies come from the
// reference by line range (see the Makefile); nothing here is derived from them.
#ifndef ORBX_SHIM_MATCH_H
#define ORBX_SHIM_MATCH_H
#include "frame_shim.h"

namespace ORB_SLAM3 {

class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);
    int SearchByProjection(Frame &F, const std::vector<MapPoint *> &vpMapPoints, const float th = 3, const bool bFarPoints = false,
                           const float thFarPoints = 50.0f);
    int SearchForInitialization(Frame &F1, Frame &F2, std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize = 10);
    int SearchByBoW(KeyFrame *pKF, Frame &F, std::vector<MapPoint *> &vpMapPointMatches);
    int SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<MapPoint *> &vpMatches12);

    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;

    // protected in the reference; public here so the glue can call ComputeThreeMaxima on its own
    float RadiusByViewingCos(const float &viewCos);
    void ComputeThreeMaxima(std::vector<int> *histo, const int L, int &ind1, int &ind2, int &ind3);

protected:
    float mfNNratio;
    bool mbCheckOrientation;
};

}  // namespace ORB_SLAM3
#endif
