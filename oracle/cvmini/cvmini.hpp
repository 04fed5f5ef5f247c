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
//
// The same header also carries what src/ORBmatcher.cc, src/Frame.cc and src/MapPoint.cc need (make ref builds them into
// oracle/_ref/liborbmatcher_ref.so): typed matrices (8U / 32S / 32F / 64F), the small CV_32F algebra those files write with cv::Mat
// (product, sum, difference, transpose, scaling, norm, convertTo, Mat_<float> comma initialiser), with OpenCV's arithmetic as
// the oracle states it and tests/test_oracle_cv2.py / test_frustum_oracle.py check it against cv2: a CV_32F product of small
// matrices is evaluated in float, left to right, without contraction; cv::norm accumulates in double.
#pragma once
#include <cassert>
#include <cmath>
#include <climits>
#include <cstddef>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <vector>
#include "../orbx_oracle.h"

typedef unsigned char uchar;
#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6

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
template <class T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;
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

inline size_t cvm_elem_size(int type) {
    switch (type) { case CV_8U: return 1; case CV_32S: case CV_32F: return 4; case CV_64F: return 8; }
    std::fprintf(stderr, "cvmini: unsupported matrix type %d\n", type); std::abort();
}
enum { NORM_L1 = 2, NORM_L2 = 4 };
struct _OutputArray;
struct Mat {
    std::shared_ptr<uchar> buf;         // owning buffer (shared by views)
    uchar *data;
    int rows, cols;
    size_t step;
    int tp;                             // CV_8U, CV_32S, CV_32F or CV_64F (single channel)
    struct Zeros { int rows, cols, type; operator Mat() const { Mat m; m = *this; return m; } };

    Mat() : data(nullptr), rows(0), cols(0), step(0), tp(CV_8U) {}
    Mat(int r, int c, int t) : data(nullptr), rows(0), cols(0), step(0), tp(CV_8U) { create(r, c, t); }
    Mat(Size s, int t) : data(nullptr), rows(0), cols(0), step(0), tp(CV_8U) { create(s.height, s.width, t); }
    // user data, not owned (cv::Mat(rows, cols, type, void *data, size_t step))
    Mat(int r, int c, int t, void *d, size_t st = 0) : data((uchar *)d), rows(r), cols(c), step(st ? st : c * cvm_elem_size(t)), tp(t) {}

    size_t elemSize() const { return cvm_elem_size(tp); }
    void create(int r, int c, int t) {
        if (data && r == rows && c == cols && t == tp) return;   // cv::Mat::create keeps a matrix of the requested shape
        const size_t e = cvm_elem_size(t);
        buf.reset((uchar *)std::malloc((size_t)std::max(r, 1) * std::max(c, 1) * e), std::free);
        data = buf.get(); rows = r; cols = c; step = (size_t)c * e; tp = t;
    }
    void release() { buf.reset(); data = nullptr; rows = cols = 0; step = 0; }
    int type() const { return tp; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t step1() const { return step / elemSize(); }
    Mat view(int y, int x, int h, int w) const {
        Mat m; m.buf = buf; m.data = data + (size_t)y * step + (size_t)x * elemSize(); m.rows = h; m.cols = w; m.step = step; m.tp = tp; return m;
    }
    Mat rowRange(int a, int b) const { return view(a, 0, b - a, cols); }
    Mat colRange(int a, int b) const { return view(0, a, rows, b - a); }
    Mat row(int y) const { return view(y, 0, 1, cols); }
    Mat col(int x) const { return view(0, x, rows, 1); }
    Mat operator()(const Rect &r) const { return view(r.y, r.x, r.height, r.width); }
    void copyTo(Mat &m) const {
        if (empty()) { m.release(); return; }
        m.create(rows, cols, tp);
        for (int y = 0; y < rows; y++) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, cols * elemSize());
    }
    inline void copyTo(const _OutputArray &o) const;
    Mat clone() const { Mat m; copyTo(m); return m; }
    inline void convertTo(const _OutputArray &o, int type) const;
    static Zeros zeros(int r, int c, int t) { Zeros z = {r, c, t}; return z; }
    Mat &operator=(const Zeros &z) {     // MatExpr assignment: create() (a no-op on a matching view), then fill
        create(z.rows, z.cols, z.type);
        for (int y = 0; y < rows; y++) std::memset(data + (size_t)y * step, 0, cols * elemSize());
        return *this;
    }
    static Mat ones(int r, int c, int t) {
        assert(t == CV_32F);
        Mat m(r, c, t);
        for (int i = 0; i < r * c; i++) reinterpret_cast<float *>(m.data)[i] = 1.f;
        return m;
    }
    static Mat eye(int r, int c, int t) {
        assert(t == CV_32F);
        Mat m; m = zeros(r, c, t);
        for (int i = 0; i < std::min(r, c); i++) m.at<float>(i, i) = 1.f;
        return m;
    }
    template <class T> T &at(int y, int x) { return *reinterpret_cast<T *>(data + (ptrdiff_t)y * (ptrdiff_t)step + x * (ptrdiff_t)sizeof(T)); }
    template <class T> const T &at(int y, int x) const { return *reinterpret_cast<const T *>(data + (ptrdiff_t)y * (ptrdiff_t)step + x * (ptrdiff_t)sizeof(T)); }
    // single index: element i of a row or column vector (of the row-major sequence in general)
    template <class T> T &at(int i) { return cols == 1 ? at<T>(i, 0) : at<T>(i / cols, i % cols); }
    template <class T> const T &at(int i) const { return cols == 1 ? at<T>(i, 0) : at<T>(i / cols, i % cols); }
    uchar *ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar *ptr(int y = 0) const { return data + (size_t)y * step; }
    template <class T> T *ptr(int y = 0) { return reinterpret_cast<T *>(data + (size_t)y * step); }
    template <class T> const T *ptr(int y = 0) const { return reinterpret_cast<const T *>(data + (size_t)y * step); }
    Mat t() const {
        assert(tp == CV_32F);
        Mat m(cols, rows, tp);
        for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) m.at<float>(x, y) = at<float>(y, x);
        return m;
    }
    double dot(const Mat &o) const {
        assert(tp == CV_32F && o.tp == CV_32F && rows == o.rows && cols == o.cols);
        double s = 0;
        for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) s += (double)at<float>(y, x) * o.at<float>(y, x);
        return s;
    }
    Mat reshape(int, int = 0) const { std::fprintf(stderr, "cvmini: Mat::reshape is not provided\n"); std::abort(); }
};

// Mat_<float>(r, c) << a, b, c ...
template <class T> struct Mat_ : Mat {
    Mat_(int r, int c) : Mat(r, c, sizeof(T) == 4 ? CV_32F : CV_64F) {}
    struct Init {
        Mat m; int i;
        Init &operator,(T v) { reinterpret_cast<T *>(m.data)[i++] = v; return *this; }
        operator Mat() const { return m; }
    };
    Init operator<<(T v) { Init it; it.m = *this; it.i = 0; it, v; return it; }
};

// CV_32F algebra.  Products of small matrices: float, left to right, every operation rounded (OpenCV's small-matrix gemm path).
inline Mat operator*(const Mat &a, const Mat &b) {
    assert(a.tp == CV_32F && b.tp == CV_32F && a.cols == b.rows);
    Mat m(a.rows, b.cols, CV_32F);
    for (int y = 0; y < a.rows; y++)
        for (int x = 0; x < b.cols; x++) {
            float s = a.at<float>(y, 0) * b.at<float>(0, x);
            for (int k = 1; k < a.cols; k++) s = s + a.at<float>(y, k) * b.at<float>(k, x);
            m.at<float>(y, x) = s;
        }
    return m;
}
template <class F> inline Mat cvm_map2(const Mat &a, const Mat &b, F f) {
    assert(a.tp == CV_32F && b.tp == CV_32F && a.rows == b.rows && a.cols == b.cols);
    Mat m(a.rows, a.cols, CV_32F);
    for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) m.at<float>(y, x) = f(a.at<float>(y, x), b.at<float>(y, x));
    return m;
}
inline Mat operator+(const Mat &a, const Mat &b) { return cvm_map2(a, b, [](float p, float q) { return p + q; }); }
inline Mat operator-(const Mat &a, const Mat &b) { return cvm_map2(a, b, [](float p, float q) { return p - q; }); }
inline Mat operator*(const Mat &a, double s) { const float f = (float)s; return cvm_map2(a, a, [f](float p, float) { return p * f; }); }
inline Mat operator*(double s, const Mat &a) { return a * s; }
inline Mat operator/(const Mat &a, double s) { return a * (1. / s); }
inline Mat operator-(const Mat &a) { return a * -1.0; }
inline double norm(const Mat &a, int kind = NORM_L2) {
    assert(a.tp == CV_32F);
    double s = 0;
    for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) { const double v = a.at<float>(y, x); s += kind == NORM_L1 ? std::fabs(v) : v * v; }
    return kind == NORM_L1 ? s : std::sqrt(s);
}
inline double norm(const Mat &a, const Mat &b, int kind = NORM_L2) {
    assert(a.tp == CV_32F && b.tp == CV_32F && a.rows == b.rows && a.cols == b.cols);
    double s = 0;
    for (int y = 0; y < a.rows; y++)
        for (int x = 0; x < a.cols; x++) { const double v = (double)a.at<float>(y, x) - (double)b.at<float>(y, x); s += kind == NORM_L1 ? std::fabs(v) : v * v; }
    return kind == NORM_L1 ? s : std::sqrt(s);
}

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
    _OutputArray(const Mat &a) : m(const_cast<Mat *>(&a)) {}   // OpenCV has it too: lets a temporary view be a destination
    void create(int r, int c, int t) const { m->create(r, c, t); }
    void create(Size s, int t) const { m->create(s.height, s.width, t); }
    void release() const { if (m) m->release(); }
    Mat getMat() const { return *m; }
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;
inline InputArray noArray() { static _InputArray a; return a; }

inline void Mat::copyTo(const _OutputArray &o) const { copyTo(*o.m); }
inline void Mat::convertTo(const _OutputArray &o, int type) const {
    const Mat src = *this;                              // the destination may be this very header (IL.convertTo(IL, CV_32F))
    assert(src.tp == CV_8U && type == CV_32F);
    Mat dst(src.rows, src.cols, type);
    for (int y = 0; y < src.rows; y++) for (int x = 0; x < src.cols; x++) dst.at<float>(y, x) = (float)src.at<uchar>(y, x);
    *o.m = dst;
}
inline std::ostream &operator<<(std::ostream &os, const Mat &m) {      // debug prints in the reference; the format is not OpenCV's
    os << "[";
    for (int y = 0; y < m.rows; y++) for (int x = 0; x < m.cols; x++) os << (m.tp == CV_32F ? m.at<float>(y, x) : (float)m.at<uchar>(y, x)) << (x + 1 < m.cols ? ", " : y + 1 < m.rows ? ";\n " : "");
    return os << "]";
}
inline void undistortPoints(InputArray, OutputArray, InputArray, InputArray, InputArray, InputArray) {
    std::fprintf(stderr, "cvmini: cv::undistortPoints is not provided (use zero distortion)\n"); std::abort();
}

// parsed by DBoW2's TemplatedVocabulary.h (save / load), never called by the pinned path
struct FileNode {
    FileNode operator[](const char *) const { return *this; }
    FileNode operator[](const std::string &) const { return *this; }
    FileNode operator[](int) const { return *this; }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator float() const { return 0.f; }
    operator double() const { return 0.; }
    operator std::string() const { return std::string(); }
    template <class T> void operator>>(T &) const {}
};
struct FileStorage {
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string &, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char *) const { return FileNode(); }
    FileNode operator[](const std::string &) const { return FileNode(); }
    template <class T> FileStorage &operator<<(const T &) { return *this; }
};

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
    const int cap = std::max(img.rows * img.cols / 4, 16);             // survivors of the 3x3 suppression: at most one per 2x2 block
    int *xs = static_cast<int *>(std::malloc(sizeof(int) * 3 * cap)), *ys = xs + cap, *sc = ys + cap;   // scratch outside operator new
    const int n = orbo_fast9(img.data, img.cols, img.rows, (int)img.step, threshold, xs, ys, sc, cap);
    assert(n <= cap);
    kps.reserve(n);
    for (int i = 0; i < n; i++) kps.push_back(KeyPoint((float)xs[i], (float)ys[i], 7.f, -1, (float)sc[i]));
    std::free(xs);
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
