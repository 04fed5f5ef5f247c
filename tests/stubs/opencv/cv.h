// Minimal stand-in for <opencv/cv.h>: just enough of cv::Mat / cv::KeyPoint / cv::InputArray / cv::OutputArray for
// `g++ -fsyntax-only` on active-orb-slam2_b200/adapter/*.cc against the reference's own headers in a container without
// OpenCV (tests/test_adapter_compiles.py).  Test infrastructure only; layouts that the adapter relies on (cv::KeyPoint is
// 28 bytes) are the real ones.
#pragma once
#include <cassert>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
#include <cmath>
#include <cstdlib>
#include <sstream>
#include <fstream>
#include <iostream>
#include <algorithm>
#include <numeric>
#include <limits>
#include <map>
#include <set>
#include <list>
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
namespace cv {
template <class T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; };
struct Mat {
    unsigned char *data = nullptr;
    int rows = 0, cols = 0;
    size_t step = 0;
    Mat() {}
    Mat(int, int, int) {}
    void create(int, int, int) {}
    int type() const { return 0; }
    bool empty() const { return !data; }
    Mat rowRange(int, int) const { return *this; }
    Mat colRange(int, int) const { return *this; }
    Mat row(int) const { return *this; }
    Mat clone() const { return *this; }
    Mat col(int) const { return *this; }
    Mat t() const { return *this; }
    double dot(const Mat &) const { return 0.0; }
    void copyTo(struct _OutputArray) const;
    template <class T> T &at(int, int = 0) { return *reinterpret_cast<T *>(data); }
    template <class T> const T &at(int, int = 0) const { return *reinterpret_cast<const T *>(data); }
    template <class T> T *ptr(int = 0) { return reinterpret_cast<T *>(data); }
    template <class T> const T *ptr(int = 0) const { return reinterpret_cast<const T *>(data); }
};
struct _InputArray { _InputArray() {} _InputArray(const Mat &) {} bool empty() const { return true; } Mat getMat() const { return Mat(); } };
struct _OutputArray { _OutputArray() {} _OutputArray(Mat &) {} void release() const {} };
inline Mat operator*(const Mat &a, const Mat &) { return a; }
inline Mat operator*(double, const Mat &a) { return a; }
inline Mat operator*(const Mat &a, double) { return a; }
inline Mat operator/(const Mat &a, double) { return a; }
inline Mat operator+(const Mat &a, const Mat &) { return a; }
inline Mat operator-(const Mat &a, const Mat &) { return a; }
inline Mat operator-(const Mat &a) { return a; }
inline double norm(const Mat &) { return 0.0; }
struct FileNode {
    FileNode operator[](const char *) const { return *this; }
    FileNode operator[](const std::string &) const { return *this; }
    FileNode operator[](int) const { return *this; }
    size_t size() const { return 0; }
    template <class T> operator T() const { return T(); }
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
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;
}
