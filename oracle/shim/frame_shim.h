// Stand-in for the handful of Frame / MapPoint members that the matcher functions on the hot path touch, with the reference's
// names and types (include/Frame.h:52-53,118-120,258-260,295-298,337-352; include/MapPoint.h:123,132,146,171-179).
// TEST INFRASTRUCTURE, NOT PRODUCT: used to compile line ranges of the reference's src/ORBmatcher.cc / src/Frame.cc
// (oracle/Makefile, _ref/libref_match.so) and to compile-check the drop-in include/ORBmatcher.h (tests/cpp).
#ifndef ORBX_SHIM_FRAME_H
#define ORBX_SHIM_FRAME_H
#include <climits>
#include <cmath>
#include <map>
#include <set>
#include <utility>
#include <vector>

#include "opencv2/opencv.hpp"

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

namespace cv {
struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(0) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
};
}  // namespace cv

using namespace std;  // include/Frame.h:47

// DBoW2::FeatureVector is a std::map<NodeId, std::vector<unsigned int>> (Thirdparty/DBoW2/DBoW2/FeatureVector.h:23-24); SearchByBoW only
// walks it (begin / end / lower_bound, ->first, ->second)
namespace DBoW2 {
class FeatureVector : public std::map<unsigned int, std::vector<unsigned int>> {};
}  // namespace DBoW2

namespace ORB_SLAM3 {

class GeometricCamera;   // only ever compared against null on this path (mpCamera2)

class MapPoint {
public:
    float mTrackProjX = 0, mTrackProjY = 0, mTrackDepth = 0, mTrackDepthR = 0, mTrackProjXR = 0, mTrackProjYR = 0;
    bool mbTrackInView = false, mbTrackInViewR = false;
    int mnTrackScaleLevel = 0, mnTrackScaleLevelR = -1;
    float mTrackViewCos = 0, mTrackViewCosR = 0;
    int Observations() { return nObs; }
    bool isBad() { return bad; }
    cv::Mat GetDescriptor() { return desc; }
    // shim state
    int nObs = 0;
    bool bad = false;
    cv::Mat desc;
};

class Frame {
public:
    bool PosInGrid(const cv::KeyPoint &kp, int &posX, int &posY);
    vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel = -1, const int maxLevel = -1,
                                     const bool bRight = false) const;
    void AssignFeaturesToGrid();
    void RefStereoTail(const std::vector<cv::DMatch> &matches);  // wrapper around src/Frame.cc:862-914 (see ref_match_glue.cpp)

    int N = 0;
    int Nleft = -1, Nright = -1;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors, mDescriptorsRight;
    std::vector<MapPoint *> mvpMapPoints;
    std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
    vector<float> mvScaleFactors;
    float mbf = 0, mb = 0;
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;  // static in the reference (one camera per process)
    float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0;         // static in the reference
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    std::vector<std::size_t> mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    DBoW2::FeatureVector mFeatVec;                                 // include/Frame.h:215
    GeometricCamera *mpCamera2 = nullptr;                          // include/Frame.h:321
};

// the KeyFrame members SearchByBoW(KeyFrame*, Frame&, …) touches (include/KeyFrame.h:218,331-340,366,377-378,391; src/ORBmatcher.cc:222-425)
class KeyFrame {
public:
    std::vector<MapPoint *> GetMapPointMatches() { return mps; }
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mDescriptors;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    int NLeft = -1, NRight = -1;
    GeometricCamera *mpCamera2 = nullptr;
    // shim state
    std::vector<MapPoint *> mps;
};

}  // namespace ORB_SLAM3
#endif
