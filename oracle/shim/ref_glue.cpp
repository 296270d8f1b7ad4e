// C-ABI glue around the reference's own ORB_SLAM3::ORBextractor (compiled unmodified from
// /root/reference/src/ORBextractor.cc against the shim).  TEST INFRASTRUCTURE, NOT PRODUCT.
#include <vector>

#include "ORBextractor.h"   // the reference's header (-I/root/reference/include)

extern "C" {

void *ref_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
    return new ORB_SLAM3::ORBextractor(nfeatures, scale_factor, nlevels, ini_th, min_th);
}
void ref_destroy(void *h) { delete (ORB_SLAM3::ORBextractor *)h; }

int ref_extract(void *h, const uint8_t *img, int rows, int cols, size_t step, const int32_t *rects, int n_rects,
                int lap0, int lap1, orc_keypoint *kps, uint8_t *desc, int cap, int *n_out, int *mono_index) {
    ORB_SLAM3::ORBextractor *ex = (ORB_SLAM3::ORBextractor *)h;
    cv::Mat image = (img && rows > 0 && cols > 0) ? cv::Mat(rows, cols, CV_8UC1, (void *)img, step) : cv::Mat();
    ex->mvDynamicArea.clear();
    for (int i = 0; i < n_rects; ++i) ex->mvDynamicArea.push_back(cv::Rect2i(rects[4 * i], rects[4 * i + 1], rects[4 * i + 2], rects[4 * i + 3]));
    std::vector<cv::KeyPoint> keys;
    cv::Mat descriptors, mask;
    std::vector<int> lap = {lap0, lap1};
    const int mono = (*ex)(image, mask, keys, descriptors, lap);
    if (mono_index) *mono_index = mono;
    if (mono < 0) { if (n_out) *n_out = 0; return -1; }
    const int n = (int)keys.size();
    if (n_out) *n_out = n;
    if (n > cap) return -2;
    for (int i = 0; i < n; ++i) {
        memcpy(&kps[i], &keys[i], sizeof(orc_keypoint));
        memcpy(desc + (size_t)i * 32, descriptors.ptr(i), 32);
    }
    return 0;
}

int ref_level(void *h, int level, uint8_t *dst, size_t dst_step, int *w, int *hh) {
    ORB_SLAM3::ORBextractor *ex = (ORB_SLAM3::ORBextractor *)h;
    if (level < 0 || level >= (int)ex->mvImagePyramid.size()) return -1;
    const cv::Mat &m = ex->mvImagePyramid[level];
    if (w) *w = m.cols;
    if (hh) *hh = m.rows;
    if (dst) for (int y = 0; y < m.rows; ++y) memcpy(dst + (size_t)y * dst_step, m.ptr(y), m.cols);
    return 0;
}

}
