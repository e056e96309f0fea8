/*
 * oracle/cvshim/cvshim.hpp -- TEST INFRASTRUCTURE ONLY.
 *
 * A minimal stand-in for the handful of OpenCV types and functions that the reference's
 * Core/MAGESLAM/Source/Image/OpenCVModified.cpp touches, so that THAT FILE can be compiled UNMODIFIED where it lies
 * (oracle/Makefile, target _ref/liborb_ref.so) and executed as the checker of the restated oracle and of the CUDA path.
 * OpenCV itself is not vendored in /root/reference and not installed here, which is why this exists.
 *
 * What is ours and what is the reference's:
 *   - every line of control flow and arithmetic of FAST, cornerScore, NMS, RunByImageBorder, the per-level budget,
 *     RetainBestFeatures, AdaptiveNonMaximalSuppresion, ICAngles, the key-point rescale, both descriptor loops and the
 *     DetectAndCompute orchestration runs from the reference's own source (including its CV_SSE2 branches);
 *   - cv::resize, cv::GaussianBlur and cv::fastAtan2 (un-vendored OpenCV arithmetic) forward to the restatements in
 *     orb_oracle.cpp that tests/test_oracle_vs_cv2.py pins bit-for-bit against OpenCV 4.13;
 *   - Mat / Rect / Point / Size / KeyPoint / RNG below are plain re-implementations of the documented OpenCV semantics
 *     (ROI views share the parent's data and step; Point2f -> Point conversion saturates through cvRound; RNG is the
 *     multiply-with-carry generator with the 4164903690 multiplier).
 *
 * Nothing under mageslam_b200/ includes this.
 */
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

#if defined(__SSE2__) && !defined(CVSHIM_NO_SSE2)
#include <emmintrin.h>
#define CV_SSE2 1
#else
#define CV_SSE2 0
#endif
#define CV_NEON 0

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_PI 3.1415926535897932384626433832795
#define CV_Assert(expr) do { if (!(expr)) throw cv::Exception(#expr, __FILE__, __LINE__); } while (0)

/* restated, cv2-pinned primitives (orb_oracle.cpp) */
extern "C" {
void  orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
int   orc_gaussian_blur_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int ksize);
int   orc_gaussian_blur_submatrix_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int ksize, int fused);
float orc_fast_atan2(float y, float x);
}

/* blur arithmetic selector of the shim (set by the C wrapper): 0 = what OpenCV 4.13 does (fixed point for a whole-buffer
 * source, fused float path for a submatrix), 1 = float fused everywhere, 2 = float unfused everywhere, 3 = fixed point everywhere */
extern int g_cvshim_blur_mode;

inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }

namespace cv {

struct Exception {
    const char* expr; const char* file; int line;
    Exception(const char* e, const char* f, int l) : expr(e), file(f), line(l) {}
    const char* what() const { return expr; }
};

template <typename T> inline T saturate_cast(float v) { return (T)v; }
template <> inline int saturate_cast<int>(float v) { return cvRound(v); }
template <typename T> inline T saturate_cast(double v) { return (T)v; }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }
template <typename T> inline T saturate_cast(int v) { return (T)v; }

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U> operator Point_<U>() const { return Point_<U>(saturate_cast<U>(x), saturate_cast<U>(y)); }
};
template <typename T> inline Point_<T>& operator*=(Point_<T>& a, float b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> inline Point_<T>& operator*=(Point_<T>& a, double b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> inline Point_<T>& operator*=(Point_<T>& a, int b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
    bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size_& o) const { return !(*this == o); }
};
typedef Size_<int> Size;

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T _x, T _y, T w, T h) : x(_x), y(_y), width(w), height(h) {}
    Rect_(const Point_<T>& a, const Point_<T>& b)
    {
        x = std::min(a.x, b.x); y = std::min(a.y, b.y);
        width = std::max(a.x, b.x) - x; height = std::max(a.y, b.y) - y;
    }
    bool contains(const Point_<T>& p) const { return x <= p.x && p.x < x + width && y <= p.y && p.y < y + height; }
};
typedef Rect_<int> Rect;

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct MatStep {
    size_t v;
    MatStep() : v(0) {}
    MatStep(size_t s) : v(s) {}
    operator size_t() const { return v; }
};

/* single-channel 8-bit matrix view; never owns external data, owns its own allocation through a shared vector */
struct Mat {
    enum { AUTO_STEP = 0, CONTINUOUS_FLAG = 1 << 14 };
    int rows, cols;
    uchar* data;
    MatStep step;
    /* the parent buffer this view was cut from (what Mat::locateROI reports) */
    uchar* datastart;
    int wholeRows, wholeCols;

    Mat() : rows(0), cols(0), data(nullptr), step(0), datastart(nullptr), wholeRows(0), wholeCols(0) {}
    Mat(Size sz, int type, void* d, size_t s = AUTO_STEP) { init(sz.height, sz.width, type, d, s); }
    Mat(int r, int c, int type, void* d, size_t s = AUTO_STEP) { init(r, c, type, d, s); }

    void init(int r, int c, int type, void* d, size_t s)
    {
        CV_Assert(type == CV_8UC1);
        rows = r; cols = c; data = (uchar*)d; step = MatStep(s == AUTO_STEP ? (size_t)c : s);
        datastart = data; wholeRows = r; wholeCols = c;
    }
    int type() const { return CV_8UC1; }
    int depth() const { return CV_8U; }
    int channels() const { return 1; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    size_t step1() const { return step.v; }
    bool isSubmatrix() const { return rows != wholeRows || cols != wholeCols; }
    bool isContinuous() const { return rows == 1 || step.v == (size_t)cols; }
    template <typename T> T& at(int r, int c) { return *(T*)(data + (size_t)r * step.v + c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *(const T*)(data + (size_t)r * step.v + c * sizeof(T)); }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step.v); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step.v); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step.v; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step.v; }

    Mat operator()(const Rect& r) const
    {
        CV_Assert(0 <= r.x && 0 <= r.width && r.x + r.width <= cols && 0 <= r.y && 0 <= r.height && r.y + r.height <= rows);
        Mat m(*this);
        m.rows = r.height; m.cols = r.width;
        m.data = data + (size_t)r.y * step.v + r.x;
        return m;
    }
    /* like cv::Mat::copyTo into an already allocated destination of the same size (the only use in the reference) */
    void copyTo(Mat& dst) const
    {
        CV_Assert(dst.data != nullptr && dst.rows == rows && dst.cols == cols);
        for (int y = 0; y < rows; ++y) memcpy(dst.ptr(y), ptr(y), (size_t)cols);
    }
};

enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };

inline float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

inline void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR)
{
    CV_Assert(interpolation == INTER_LINEAR && fx == 0 && fy == 0);
    CV_Assert(dst.data != nullptr && dst.size() == dsize);          /* cv::Mat::create is a no-op for a matching view */
    orc_resize_linear_u8(src.data, src.cols, src.rows, (int)src.step.v, dst.data, dst.cols, dst.rows, (int)dst.step.v);
}

/* In place on a level view, as the reference calls it (OpenCVModified.cpp:863). Border pixels are taken by REFLECT_101 of the
 * view itself; real OpenCV would read the parent buffer's pixels around a submatrix instead. That only affects output pixels
 * within ksize/2 of a level edge, which no descriptor of the tier configurations reads. */
inline void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT)
{
    CV_Assert(ksize.width == ksize.height && sigmaX == 2 && sigmaY == 2 && borderType == BORDER_REFLECT_101);
    CV_Assert(dst.data != nullptr && dst.size() == src.size());
    std::vector<uchar> tmp((size_t)src.rows * src.cols);
    int mode = g_cvshim_blur_mode;
    bool fixedPoint = mode == 3 || (mode == 0 && !src.isSubmatrix());
    int rc;
    if (fixedPoint)
        rc = orc_gaussian_blur_u8(src.data, src.cols, src.rows, (int)src.step.v, tmp.data(), src.cols, ksize.width);
    else
        rc = orc_gaussian_blur_submatrix_u8(src.data, src.cols, src.rows, (int)src.step.v, tmp.data(), src.cols, ksize.width, mode == 2 ? 0 : 1);
    CV_Assert(rc == 0);
    for (int y = 0; y < src.rows; ++y) memcpy(dst.ptr(y), tmp.data() + (size_t)y * src.cols, (size_t)src.cols);
}

/* cv::RNG: multiply-with-carry, state = (unsigned)state * 4164903690U + (state >> 32) */
struct RNG {
    uint64_t state;
    RNG() : state(0xffffffff) {}
    RNG(uint64_t s) : state(s ? s : 0xffffffff) {}
    unsigned next() { state = (uint64_t)(unsigned)state * 4164903690U + (unsigned)(state >> 32); return (unsigned)state; }
    int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

template <typename T> inline T* alignPtr(T* ptr, int n = (int)sizeof(T)) { return (T*)(((size_t)ptr + n - 1) & -(size_t)n); }

} /* namespace cv */
