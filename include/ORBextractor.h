/* ORBextractor.h — drop-in replacement for /root/reference/include/ORBextractor.h (+ src/ORBextractor.cc).
 *
 * Same namespace, class name and public signatures as the reference (include/ORBextractor.h:43-110), so
 * Frame, Tracking, LocalMapping and LoopClosing compile unchanged; every call forwards to the sm_100a CUDA
 * library through the C ABI of orbx.h.  Header-only: put this directory before the reference's include/
 * on the include path, drop src/ORBextractor.cc from the build and link liborbx.so.
 *
 * Differences that a maintainer should know about:
 *  - the protected helpers of the reference (ComputePyramid, DistributeOctTree, ...) do not exist here: the
 *    work happens on the GPU; ExtractorNode is kept as a plain type for source compatibility only;
 *  - mvImagePyramid is NOT refreshed on every call (nothing in the reference reads it outside the extractor):
 *    call MaterializePyramid() — or set mbMaterializePyramid — to copy the levels back, with the reference's
 *    19-px REFLECT_101 border, as ROI views exactly like src/ORBextractor.cc:1216-1217;
 *  - -1 is returned for an empty image only, as in the reference; a missing CUDA device, a failed allocation or a CUDA error
 *    throws std::runtime_error (there is no CPU fallback, and a dead GPU must not look like a blank frame);
 *  - ORBX_DEVICE (environment) selects the CUDA device, default 0; the handle is created for the first
 *    image size it sees and re-created if a larger image arrives.
 */
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <cstdlib>
#include <iostream>
#include <list>
#include <stdexcept>
#include <string>
#include <vector>

#include <opencv2/opencv.hpp>

#include "orbx.h"

namespace ORB_SLAM3
{

class ExtractorNode
{
public:
    ExtractorNode():bNoMore(false){}
    std::vector<cv::KeyPoint> vKeys;
    cv::Point2i UL, UR, BL, BR;
    std::list<ExtractorNode>::iterator lit;
    bool bNoMore;
};

class ORBextractor
{
public:

    enum {HARRIS_SCORE=0, FAST_SCORE=1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
        : mbMaterializePyramid(false), nfeatures(nfeatures), scaleFactor(scaleFactor), nlevels(nlevels),
          iniThFAST(iniThFAST), minThFAST(minThFAST), mpHandle(nullptr), mMaxW(0), mMaxH(0)
    {
        const char* dev = std::getenv("ORBX_DEVICE");
        mDevice = dev ? std::atoi(dev) : 0;
        mvImagePyramid.resize(nlevels);
        if(!EnsureHandle(640, 480))
            Fail("orbx_create");
    }

    ~ORBextractor(){ if(mpHandle) orbx_destroy(mpHandle); }

    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image (src/ORBextractor.cc:1125-1207).
    // Mask is ignored, as in the reference.
    int operator()( cv::InputArray _image, cv::InputArray _mask,
                    std::vector<cv::KeyPoint>& _keypoints,
                    cv::OutputArray _descriptors, std::vector<int> &vLappingArea)
    {
        (void)_mask;
        if(_image.empty())
            return -1;
        cv::Mat image = _image.getMat();
        assert(image.type() == CV_8UC1 );
        if(!EnsureHandle(image.cols, image.rows))
            return Fail("orbx_create");

        std::vector<int32_t> rects;
        rects.reserve(4*mvDynamicArea.size());
        for(const cv::Rect2i& r : mvDynamicArea)
        {
            rects.push_back(r.x); rects.push_back(r.y); rects.push_back(r.width); rects.push_back(r.height);
        }
        const int cap = nfeatures + 8*nlevels + 64;      // the quadtree can overshoot a level's quota by ≤ 2
        mKeyBuf.resize(cap);
        mDescBuf.resize((size_t)cap*32);
        int n = 0, monoIndex = -1;
        const int lap0 = vLappingArea.size()>0 ? vLappingArea[0] : 0, lap1 = vLappingArea.size()>1 ? vLappingArea[1] : 0;
        const int rc = orbx_extract(mpHandle, image.data, image.rows, image.cols, (size_t)image.step,
                                    rects.empty() ? nullptr : rects.data(), (int)mvDynamicArea.size(), lap0, lap1,
                                    mKeyBuf.data(), mDescBuf.data(), cap, &n, &monoIndex);
        if(rc == ORBX_EMPTY)
            return -1;
        if(rc != ORBX_OK)
            return Fail("orbx_extract");

        if( n == 0 )
            _descriptors.release();
        else
        {
            _descriptors.create(n, 32, CV_8U);
            cv::Mat descriptors = _descriptors.getMat();
            for(int i=0; i<n; i++)
                memcpy(descriptors.ptr(i), &mDescBuf[(size_t)i*32], 32);
        }
        _keypoints = std::vector<cv::KeyPoint>(n);
        static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "cv::KeyPoint layout");
        if(n > 0)
            memcpy((void*)_keypoints.data(), mKeyBuf.data(), (size_t)n*sizeof(orbx_keypoint));
        if(mbMaterializePyramid)
            MaterializePyramid();
        return monoIndex;
    }

    int inline GetLevels(){
        return nlevels;}

    float inline GetScaleFactor(){
        return scaleFactor;}

    std::vector<float> inline GetScaleFactors(){
        return Param(0);
    }

    std::vector<float> inline GetInverseScaleFactors(){
        return Param(1);
    }

    std::vector<float> inline GetScaleSigmaSquares(){
        return Param(2);
    }

    std::vector<float> inline GetInverseScaleSigmaSquares(){
        return Param(3);
    }

    // Copies the pyramid of the last call back from the GPU into mvImagePyramid (padded planes, ROI views).
    void MaterializePyramid()
    {
        for(int level=0; level<nlevels; level++)
        {
            int w=0, h=0;
            if(orbx_level_size(mpHandle, level, &w, &h) != ORBX_OK)
                return;
            cv::Mat temp(h + 38, w + 38, CV_8UC1);
            if(orbx_get_pyramid(mpHandle, 0, level, 1, temp.data, (size_t)temp.step) != ORBX_OK)
                return;
            mvImagePyramid[level] = temp(cv::Rect(19, 19, w, h));
        }
    }

    // The C-ABI handle (for orbx_stereo_matches, which reads the two extractors' device-resident pyramids, and for batch calls).
    orbx_extractor* Handle(){ return mpHandle; }

    std::vector<cv::Mat> mvImagePyramid;
    std::vector<cv::Rect2i> mvDynamicArea;
    bool mbMaterializePyramid;

protected:

    bool EnsureHandle(int w, int h)
    {
        if(mpHandle && w <= mMaxW && h <= mMaxH)
            return true;
        if(mpHandle) orbx_destroy(mpHandle);
        mMaxW = std::max(w, mMaxW); mMaxH = std::max(h, mMaxH);
        mpHandle = orbx_create(nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST, mDevice, mMaxW, mMaxH, 1);
        if(!mpHandle)
            std::cerr << "[ORBextractor/orbx] " << orbx_last_error(nullptr) << std::endl;
        return mpHandle != nullptr;
    }

    int Fail(const char* what)
    {
        // -1 means "empty image" to the callers (src/ORBextractor.cc:1129): a missing device, a failed allocation or a CUDA error
        // must not look like a blank frame.  The reference has no recoverable error path either, so this throws.
        const std::string msg = std::string("[ORBextractor/orbx] ") + what + " failed: " + orbx_last_error(mpHandle);
        std::cerr << msg << std::endl;
        throw std::runtime_error(msg);
    }

    std::vector<float> Param(int which)
    {
        if(!mpHandle)
            Fail("orbx_params (no device handle)");
        std::vector<float> v[4];
        for(auto& x : v) x.resize(nlevels);
        orbx_params(mpHandle, v[0].data(), v[1].data(), v[2].data(), v[3].data(), nullptr);
        return v[which];
    }

    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;

    orbx_extractor* mpHandle;
    int mDevice, mMaxW, mMaxH;
    std::vector<orbx_keypoint> mKeyBuf;
    std::vector<uint8_t> mDescBuf;
};

} //namespace ORB_SLAM

#endif
