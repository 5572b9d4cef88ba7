/*
 * oracle/cvshim/opencv2/opencv.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A minimal stand-in for the OpenCV C++ API, just large enough to compile the reference's own
 *   /root/reference/src/slam_system/orb_slam3/src/ORBextractor.cc
 * where it lies (OpenCV's C++ headers are not installed in the build container).  Containers
 * (Mat, KeyPoint, Point_, ...) are re-implemented here for CV_8UC1 only; the four image primitives
 * (FAST, resize, GaussianBlur, fastAtan2) forward to the C models in oracle/cvmodels.c, which are
 * pinned bit-exactly against the real cv2 4.13.0 (tests/test_oracle_cv2.py).  The result,
 * oracle/_ref/libref_orb.so, is the reference's in-tree extractor logic -- octree with libstdc++'s
 * std::sort and std::list, IC_Angle, rotated BRIEF with libm's cosf/sinf, the output ordering --
 * compiled by gcc, and is what oracle/orb_oracle.cpp is checked against (tests/test_ref_build.py).
 */
#ifndef DVM_CVSHIM_OPENCV_HPP
#define DVM_CVSHIM_OPENCV_HPP

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "cvmodels.h"

typedef unsigned char uchar;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0

inline int cvRound(double v) { return cvm_round_d(v); }
inline int cvRound(float v) { return cvm_round_f(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }

namespace cv {

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) { }
    Point_(T x_, T y_) : x(x_), y(y_) { }
    template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) { }
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;

struct Size {
    int width, height;
    Size() : width(0), height(0) { }
    Size(int w, int h) : width(w), height(h) { }
};
struct Rect {
    int x, y, width, height;
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) { }
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) { }
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) { }
};

/* CV_8UC1 matrix header over shared storage (ROI views share the buffer, like cv::Mat) */
class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    uchar* data = nullptr;
    Mat() { }
    Mat(int r, int c, int /*type*/) { create(r, c, 0); }
    Mat(Size s, int /*type*/) { create(s.height, s.width, 0); }
    void create(int r, int c, int /*type*/)
    {
        if (r == rows && c == cols && data && buf_ && step == (size_t)c) return;
        rows = r; cols = c; step = (size_t)c;
        buf_ = std::shared_ptr<std::vector<uchar>>(new std::vector<uchar>((size_t)r * c));
        data = buf_->data();
    }
    static Mat zeros(int r, int c, int t) { Mat m(r, c, t); std::fill(m.buf_->begin(), m.buf_->end(), 0); return m; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    size_t step1() const { return step; }
    void release() { rows = cols = 0; step = 0; data = nullptr; buf_.reset(); }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + x); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + x); }
    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
    Mat operator()(const Rect& r) const { Mat m = *this; m.data = data + (size_t)r.y * step + r.x; m.rows = r.height; m.cols = r.width; return m; }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
    Mat row(int y) const { return (*this)(Rect(0, y, cols, 1)); }
    Mat clone() const
    {
        Mat m(rows, cols, 0);
        for (int y = 0; y < rows; y++) memcpy(m.ptr(y), ptr(y), cols);
        return m;
    }
    void copyTo(Mat dst) const
    {
        assert(dst.rows == rows && dst.cols == cols);
        for (int y = 0; y < rows; y++) memcpy(dst.ptr(y), ptr(y), cols);
    }
private:
    std::shared_ptr<std::vector<uchar>> buf_;
};

/* InputArray / OutputArray: thin references to a Mat */
class _InputArray {
public:
    _InputArray() : m_(nullptr) { }
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) { }
    bool empty() const { return !m_ || m_->empty(); }
    Mat getMat() const { return m_ ? *m_ : Mat(); }
protected:
    Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray(Mat& m) { m_ = &m; }
    void release() const { m_->release(); }
    void create(int r, int c, int t) const { m_->create(r, c, t); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

/* ---- the four external primitives, forwarded to the cv2-pinned C models ---- */
inline void FAST(const Mat& img, std::vector<KeyPoint>& kps, int threshold, bool nonmaxSuppression = true)
{
    assert(nonmaxSuppression);
    (void)nonmaxSuppression;
    std::vector<cvm_fast_kp> buf((size_t)img.rows * img.cols / 4 + 16);
    int n = cvm_fast_detect(img.data, img.cols, img.rows, (int)img.step, threshold, buf.data(), (int)buf.size());
    kps.clear();
    for (int i = 0; i < n; i++) kps.push_back(KeyPoint((float)buf[i].x, (float)buf[i].y, 7.f, -1.f, (float)buf[i].response));
}
inline void resize(const Mat& src, Mat& dst, Size dsize, double, double, int interpolation)
{
    assert(interpolation == INTER_LINEAR && dst.rows == dsize.height && dst.cols == dsize.width);
    (void)interpolation;
    cvm_resize_linear_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, dsize.width, dsize.height, (int)dst.step);
}
inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int /*borderType: REFLECT_101*/)
{
    assert(dst.rows == src.rows + top + bottom && dst.cols == src.cols + left + right);
    auto refl = [](int p, int n) { if (n == 1) return 0; while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p; return p; };
    Mat s = src.clone(); // src may be a view into dst (the reference pads the pyramid level in place)
    for (int y = 0; y < dst.rows; y++) {
        const uchar* srow = s.ptr(refl(y - top, s.rows));
        uchar* drow = dst.ptr(y);
        for (int x = 0; x < dst.cols; x++) drow[x] = srow[refl(x - left, s.cols)];
    }
}
inline void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sx, double sy, int borderType)
{
    assert(ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2 && borderType == BORDER_REFLECT_101);
    (void)ksize; (void)sx; (void)sy; (void)borderType;
    Mat tmp = src.clone();
    cvm_gaussian7_u8(tmp.data, tmp.cols, tmp.rows, (int)tmp.step, dst.data, (int)dst.step);
}
inline float fastAtan2(float y, float x) { return cvm_fast_atan2(y, x); }

struct KeyPointsFilter { /* only used by the dead ComputeKeyPointsOld */
    static void retainBest(std::vector<KeyPoint>& kps, int n)
    {
        if (n >= 0 && (int)kps.size() > n) {
            std::nth_element(kps.begin(), kps.begin() + n, kps.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
            kps.resize(n);
        }
    }
};

} // namespace cv
#endif
