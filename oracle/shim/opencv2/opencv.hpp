// Minimal stand-in for the few OpenCV types/functions that /root/reference/src/ORBextractor.cc uses, so
// that the reference's own source can be compiled UNMODIFIED in a container without OpenCV C++ headers
// (oracle/Makefile target `ref` → oracle/_ref/libref_orb.so).  TEST INFRASTRUCTURE, NOT PRODUCT.
//
// Only the container classes are implemented here; the six numeric primitives (resize, copyMakeBorder,
// FAST, GaussianBlur, fastAtan2, cvRound) forward to liborb_oracle.so's restatements, which the test
// suite pins bit-for-bit to the real cv2 4.13.0.  Nothing here is derived from OpenCV sources.
#ifndef ORBX_SHIM_OPENCV_HPP
#define ORBX_SHIM_OPENCV_HPP

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../../orb_oracle.h"

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5

typedef unsigned char uchar;

inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return orc_cvround(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }

namespace cv {

template <typename T> inline T saturate_cast(float v) { return (T)v; }
template <> inline int saturate_cast<int>(float v) { return cvRound(v); }
template <typename T> inline T saturate_cast(int v) { return (T)v; }

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U> operator Point_<U>() const { return Point_<U>(saturate_cast<U>(x), saturate_cast<U>(y)); }
    template <typename S> Point_ &operator*=(S s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

template <typename T>
struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <typename T>
struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T _x, T _y, T w, T h) : x(_x), y(_y), width(w), height(h) {}
    bool contains(const Point_<T> &p) const { return x <= p.x && p.x < x + width && y <= p.y && p.y < y + height; }
};
typedef Rect_<int> Rect2i;
typedef Rect_<int> Rect;

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};
static_assert(sizeof(KeyPoint) == 28, "KeyPoint layout");

struct MatStep {
    size_t v;
    MatStep(size_t s = 0) : v(s) {}
    operator size_t() const { return v; }
};

class Mat {
public:
    int rows, cols;
    uchar *data;
    MatStep step;
    Mat() : rows(0), cols(0), data(nullptr), step(0) {}
    Mat(int r, int c, int type) { alloc(r, c, type); }
    Mat(Size s, int /*type*/) { alloc(s.height, s.width); }
    Mat(int r, int c, int /*type*/, void *ext, size_t st) : rows(r), cols(c), data((uchar *)ext), step(st) {}
    static Mat zeros(int r, int c, int t) { Mat m(r, c, t); if (m.buf) std::fill(m.buf->begin(), m.buf->end(), 0); return m; }
    void create(int r, int c, int type) { if (r != rows || c != cols || !data) alloc(r, c, type); }
    void release() { buf.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return !data || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    size_t step1() const { return step; }
    Mat operator()(const Rect &r) const { return view(r.y, r.x, r.height, r.width); }
    Mat rowRange(int a, int b) const { return view(a, 0, b - a, cols); }
    Mat colRange(int a, int b) const { return view(0, a, rows, b - a); }
    Mat row(int i) const { return view(i, 0, 1, cols); }
    template <typename T> T &at(int y, int x) { return *(T *)(data + (size_t)y * step + x * sizeof(T)); }
    template <typename T> const T &at(int y, int x) const { return *(const T *)(data + (size_t)y * step + x * sizeof(T)); }
    template <typename T> T *ptr(int y = 0) { return (T *)(data + (size_t)y * step); }
    template <typename T> const T *ptr(int y = 0) const { return (const T *)(data + (size_t)y * step); }
    uchar *ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar *ptr(int y = 0) const { return data + (size_t)y * step; }
    Mat clone() const {
        Mat m(rows, cols, CV_8UC1);
        for (int y = 0; y < rows; ++y) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, cols);
        return m;
    }
    void copyTo(Mat dst) const {  // dst is a (row) view into preallocated storage
        for (int y = 0; y < rows; ++y) memcpy(dst.data + (size_t)y * dst.step, data + (size_t)y * step, cols);
    }

private:
    std::shared_ptr<std::vector<uchar>> buf;
    void alloc(int r, int c, int type = CV_8UC1) {
        const size_t es = type == CV_32F ? 4 : 1;   // DBoW2's FORB::toMat32F is the only non-8U user (compiled, never called)
        rows = r; cols = c; step = (size_t)c * es;
        buf = std::make_shared<std::vector<uchar>>((size_t)r * c * es + 1);
        data = buf->data();
    }
    Mat view(int y, int x, int h, int w) const {
        Mat m;
        m.rows = h; m.cols = w; m.step = step; m.buf = buf;
        m.data = data + (size_t)y * step + x;
        return m;
    }
};

class _InputArray {
public:
    _InputArray(const Mat &m) : m_(&m) {}
    bool empty() const { return m_->empty(); }
    Mat getMat() const { return *m_; }
private:
    const Mat *m_;
};
class _OutputArray {
public:
    _OutputArray(Mat &m) : m_(&m) {}
    void release() const { m_->release(); }
    void create(int r, int c, int t) const { m_->create(r, c, t); }
    Mat getMat() const { return *m_; }
private:
    Mat *m_;
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };

inline void resize(const Mat &src, Mat &dst, Size sz, double, double, int) {
    dst.create(sz.height, sz.width, CV_8UC1);  // no-op for the preallocated ROI the reference passes
    orc_resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}
inline void copyMakeBorder(const Mat &src, Mat &dst, int top, int bottom, int left, int right, int /*borderType*/) {
    std::vector<uchar> tmp((size_t)src.rows * src.cols);  // src may live inside dst (the pyramid ROI)
    for (int y = 0; y < src.rows; ++y) memcpy(&tmp[(size_t)y * src.cols], src.data + (size_t)y * src.step, src.cols);
    dst.create(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
    (void)right; (void)bottom;
    orc_border_reflect101_u8(tmp.data(), src.cols, src.rows, src.cols, dst.data, dst.step, top);
}
inline void FAST(const Mat &img, std::vector<KeyPoint> &kps, int threshold, bool /*nonmaxSuppression = true*/) {
    std::vector<int32_t> xys(3 * 4096);
    int n = orc_fast9_nms(img.data, img.cols, img.rows, img.step, threshold, xys.data(), 4096);
    if (n > 4096) { xys.resize(3 * (size_t)n); n = orc_fast9_nms(img.data, img.cols, img.rows, img.step, threshold, xys.data(), n); }
    kps.clear();
    for (int i = 0; i < n; ++i) kps.push_back(KeyPoint((float)xys[3 * i], (float)xys[3 * i + 1], 7.f, -1, (float)xys[3 * i + 2]));
}
inline void GaussianBlur(const Mat &src, Mat &dst, Size, double, double, int) {
    std::vector<uchar> out((size_t)src.rows * src.cols);
    orc_gaussian7_u8(src.data, src.cols, src.rows, src.step, out.data(), src.cols);
    for (int y = 0; y < src.rows; ++y) memcpy(dst.data + (size_t)y * dst.step, &out[(size_t)y * src.cols], src.cols);
}
inline float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

// cv::FileStorage / cv::FileNode: named by DBoW2's YAML save()/load(), which the oracle build never calls (the
// vocabulary is read with the reference's own loadFromTextFile).  These stubs only have to compile.
class FileNode {
public:
    FileNode operator[](const std::string &) const { return FileNode(); }
    FileNode operator[](const char *) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator double() const { return 0.0; }
    operator std::string() const { return std::string(); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage(const std::string &, int) {}
    bool isOpened() const { return false; }
    FileNode operator[](const std::string &) const { return FileNode(); }
    FileNode operator[](const char *) const { return FileNode(); }
};
template <typename T> inline FileStorage &operator<<(FileStorage &fs, const T &) { return fs; }

struct KeyPointsFilter {  // only referenced by the reference's dead ComputeKeyPointsOld (never called)
    static void retainBest(std::vector<KeyPoint> &, int) {}
};

}  // namespace cv
#endif
