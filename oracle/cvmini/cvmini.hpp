// cvmini — a FUNCTIONAL stand-in for the handful of OpenCV types and functions that the reference's
// src/ORBextractor.cc uses, so that the reference's own translation unit compiles UNMODIFIED in an image without
// OpenCV (oracle/Makefile target `ref`, output oracle/_ref/liborbextractor_ref.so).  TEST INFRASTRUCTURE ONLY:
// nothing in the product library includes or links this.
//
// What is real and what is stood in:
//   * everything ORB-SLAM2 wrote itself runs from the reference's source: the constructor tables, ComputePyramid's
//     level sizes, the per-cell FAST loop and its minThFAST retry, ExtractorNode::DivideNode / DistributeOctTree,
//     IC_Angle, computeOrbDescriptor (pattern, rotation, cvRound), the level concatenation and rescale;
//   * the OpenCV primitives it delegates to (cv::resize INTER_LINEAR, cv::copyMakeBorder REFLECT_101, cv::FAST with
//     non-max suppression, cv::GaussianBlur 7x7 sigma 2, cv::fastAtan2, cvRound) are the oracle's restatements
//     (oracle/extract_oracle.c), each of which tests/test_oracle_cv2.py checks bit for bit against cv2 4.13.
// cv::Mat semantics that the reference relies on are kept: reference-counted buffers, ROI views that share them,
// `create()` as a no-op on a matrix that already has the requested shape (so resize / copyMakeBorder / Mat::zeros
// write THROUGH a view into its parent buffer, ORBextractor.cc:1037,1120,1122), clone() = compact copy of the view.
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <list>
#include <memory>
#include <vector>
#include "../orbx_oracle.h"

typedef unsigned char uchar;
#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0

inline int cvRound(double v) { return (int)lrint(v); }   // round half to even, as SSE2 cvtsd2si
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {
using ::cvRound; using ::cvFloor; using ::cvCeil;

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
    template <class S> Point_ &operator*=(S s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect(int a, int b, int w, int h) : x(a), y(b), width(w), height(h) {} };

struct KeyPoint {                       // 28 bytes, as OpenCV's
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };

struct Mat {
    std::shared_ptr<uchar> buf;         // owning buffer (shared by views)
    uchar *data;
    int rows, cols;
    size_t step;
    struct Zeros { int rows, cols, type; };

    Mat() : data(nullptr), rows(0), cols(0), step(0) {}
    Mat(int r, int c, int) : data(nullptr), rows(0), cols(0), step(0) { create(r, c, CV_8UC1); }
    Mat(Size s, int) : data(nullptr), rows(0), cols(0), step(0) { create(s.height, s.width, CV_8UC1); }
    // user data, not owned (cv::Mat(rows, cols, type, void *data, size_t step))
    Mat(int r, int c, int, void *d, size_t st) : data((uchar *)d), rows(r), cols(c), step(st) {}

    void create(int r, int c, int) {
        if (data && r == rows && c == cols) return;   // cv::Mat::create keeps a matrix of the requested shape
        buf.reset((uchar *)std::malloc((size_t)std::max(r, 1) * std::max(c, 1)), std::free);
        data = buf.get(); rows = r; cols = c; step = (size_t)c;
    }
    void release() { buf.reset(); data = nullptr; rows = cols = 0; step = 0; }
    int type() const { return CV_8UC1; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t step1() const { return step; }
    Mat view(int y, int x, int h, int w) const {
        Mat m; m.buf = buf; m.data = data + (size_t)y * step + x; m.rows = h; m.cols = w; m.step = step; return m;
    }
    Mat rowRange(int a, int b) const { return view(a, 0, b - a, cols); }
    Mat colRange(int a, int b) const { return view(0, a, rows, b - a); }
    Mat operator()(const Rect &r) const { return view(r.y, r.x, r.height, r.width); }
    Mat clone() const {
        Mat m; if (empty()) return m;
        m.create(rows, cols, CV_8UC1);
        for (int y = 0; y < rows; y++) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, cols);
        return m;
    }
    static Zeros zeros(int r, int c, int t) { Zeros z = {r, c, t}; return z; }
    Mat &operator=(const Zeros &z) {     // MatExpr assignment: create() (a no-op on a matching view), then fill
        create(z.rows, z.cols, z.type);
        for (int y = 0; y < rows; y++) std::memset(data + (size_t)y * step, 0, cols);
        return *this;
    }
    template <class T> T &at(int y, int x) { return *reinterpret_cast<T *>(data + (ptrdiff_t)y * (ptrdiff_t)step + x * (ptrdiff_t)sizeof(T)); }
    template <class T> const T &at(int y, int x) const { return *reinterpret_cast<const T *>(data + (ptrdiff_t)y * (ptrdiff_t)step + x * (ptrdiff_t)sizeof(T)); }
    uchar *ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar *ptr(int y = 0) const { return data + (size_t)y * step; }
    template <class T> T *ptr(int y = 0) { return reinterpret_cast<T *>(data + (size_t)y * step); }
    template <class T> const T *ptr(int y = 0) const { return reinterpret_cast<const T *>(data + (size_t)y * step); }
};

struct _InputArray {
    const Mat *m;
    _InputArray() : m(nullptr) {}
    _InputArray(const Mat &a) : m(&a) {}
    bool empty() const { return !m || m->empty(); }
    Mat getMat() const { return m ? *m : Mat(); }
};
struct _OutputArray {
    Mat *m;
    _OutputArray() : m(nullptr) {}
    _OutputArray(Mat &a) : m(&a) {}
    void create(int r, int c, int t) const { m->create(r, c, t); }
    void create(Size s, int t) const { m->create(s.height, s.width, t); }
    void release() const { if (m) m->release(); }
    Mat getMat() const { return *m; }
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;
inline InputArray noArray() { static _InputArray a; return a; }

inline float fastAtan2(float y, float x) { return orbo_fast_atan2(y, x); }

// cv::resize(8UC1, INTER_LINEAR) with an explicit destination size
inline void resize(InputArray src_, OutputArray dst_, Size sz, double, double, int) {
    Mat src = src_.getMat();
    dst_.create(sz, CV_8UC1);
    Mat dst = dst_.getMat();
    orbo_resize_linear_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

// cv::copyMakeBorder(..., BORDER_REFLECT_101 [+ BORDER_ISOLATED]) with equal borders.  Without BORDER_ISOLATED
// OpenCV would take the border from the parent of a source VIEW; ORBextractor.cc:1127 passes the caller's whole image
// there, so both calls reflect the source's own pixels.
inline void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int) {
    Mat src = src_.getMat();
    assert(top == bottom && top == left && top == right);
    dst_.create(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
    Mat dst = dst_.getMat();
    uchar *interior = dst.data + (size_t)top * dst.step + left;
    if (interior != src.data)
        for (int y = 0; y < src.rows; y++) std::memmove(interior + (size_t)y * dst.step, src.data + (size_t)y * src.step, src.cols);
    orbo_border_reflect101(interior, src.cols, src.rows, (int)dst.step, top);
}

// cv::GaussianBlur(8U, Size(7,7), 2, 2, BORDER_REFLECT_101)
inline void GaussianBlur(InputArray src_, OutputArray dst_, Size k, double sx, double sy, int) {
    Mat src = src_.getMat();
    assert(k.width == 7 && k.height == 7 && sx == 2 && sy == 2);
    dst_.create(src.rows, src.cols, CV_8UC1);
    Mat dst = dst_.getMat();
    orbo_gaussian7_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, (int)dst.step);   // reads src fully before it writes dst
}

// cv::FAST (TYPE_9_16) with non-max suppression: row-major order, KeyPoint(x, y, 7.f, -1, score)
inline void FAST(InputArray img_, std::vector<KeyPoint> &kps, int threshold, bool nms) {
    Mat img = img_.getMat();
    assert(nms);
    kps.clear();
    const int cap = std::max(img.rows * img.cols, 1);
    std::vector<int> xs(cap), ys(cap), sc(cap);
    const int n = orbo_fast9(img.data, img.cols, img.rows, (int)img.step, threshold, xs.data(), ys.data(), sc.data(), cap);
    for (int i = 0; i < n; i++) kps.push_back(KeyPoint((float)xs[i], (float)ys[i], 7.f, -1, (float)sc[i]));
}

// only ComputeKeyPointsOld (dead code, ORBextractor.cc:855) uses this
struct KeyPointsFilter {
    static void retainBest(std::vector<KeyPoint> &k, int n) {
        if (n < 0 || (int)k.size() <= n) return;
        std::stable_sort(k.begin(), k.end(), [](const KeyPoint &a, const KeyPoint &b) { return a.response > b.response; });
        k.resize(n);
    }
};
}  // namespace cv
