// mini-cv: a minimal, from-scratch stand-in for the few OpenCV types and functions that the
// reference's ORBextractor.cc / ORBextractor.h use.  TEST INFRASTRUCTURE ONLY: it lets the
// reference's unmodified sources compile into oracle/_ref on a box with no OpenCV C++ SDK,
// and lets tests compile the drop-in ORBextractor shim.  The image primitives declared here
// are implemented in oracle/minicv/minicv_impl.cc on top of the C oracle, which is pinned
// bit-for-bit against cv2 4.13 (tests/test_oracle_primitives.py).
#ifndef MINICV_CORE_HPP
#define MINICV_CORE_HPP
#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32S 4
#define CV_32F 5

inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }

namespace cv {
typedef unsigned char uchar;

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
};
template <typename T> inline Point_<T>& operator*=(Point_<T>& a, float b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <typename T> inline Point_<T>& operator+=(Point_<T>& a, const Point_<T>& b) { a.x += b.x; a.y += b.y; return a; }
template <typename T> inline bool operator==(const Point_<T>& a, const Point_<T>& b) { return a.x == b.x && a.y == b.y; }
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T _x, T _y, T w, T h) : x(_x), y(_y), width(w), height(h) {}
};
typedef Rect_<int> Rect;

struct Range { int start, end; Range(int s, int e) : start(s), end(e) {} };

class KeyPoint {
public:
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};

struct MatZeros { int rows, cols, type; };

// 8-bit single-channel (or, for elemSize, CV_32S/CV_32F) reference-counted 2-D array.
class Mat {
public:
    int flags, rows, cols;
    uchar* data;
    size_t step;
    Mat() : flags(0), rows(0), cols(0), data(0), step(0), rc(0), owner(0) {}
    Mat(int r, int c, int type) : flags(0), rows(0), cols(0), data(0), step(0), rc(0), owner(0) { create(r, c, type); }
    Mat(Size s, int type) : flags(0), rows(0), cols(0), data(0), step(0), rc(0), owner(0) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, void* ext, size_t st = 0)
        : flags(type), rows(r), cols(c), data((uchar*)ext), step(st ? st : (size_t)c * esz(type)), rc(0), owner(0) {}
    Mat(const Mat& m) : flags(m.flags), rows(m.rows), cols(m.cols), data(m.data), step(m.step), rc(m.rc), owner(m.owner) { if (rc) ++*rc; }
    ~Mat() { release(); }
    Mat& operator=(const Mat& m) {
        if (this != &m) {
            if (m.rc) ++*m.rc;
            release();
            flags = m.flags; rows = m.rows; cols = m.cols; data = m.data; step = m.step; rc = m.rc; owner = m.owner;
        }
        return *this;
    }
    // "m = Mat::zeros(r,c,t)": like OpenCV's MatExpr assignment this re-uses the existing
    // buffer when size and type already match (ORBextractor.cc:1057 relies on that).
    Mat& operator=(const MatZeros& z) {
        create(z.rows, z.cols, z.type);
        for (int y = 0; y < rows; y++) memset(data + (size_t)y * step, 0, (size_t)cols * esz(flags));
        return *this;
    }
    Mat(const MatZeros& z) : flags(0), rows(0), cols(0), data(0), step(0), rc(0), owner(0) { *this = z; }
    static MatZeros zeros(int r, int c, int type) { MatZeros z = {r, c, type}; return z; }
    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == flags) return;
        release();
        flags = type; rows = r; cols = c; step = (size_t)c * esz(type);
        size_t bytes = step * (size_t)r;
        owner = (uchar*)std::malloc(bytes ? bytes : 1);
        data = owner;
        rc = (int*)std::malloc(sizeof(int));
        *rc = 1;
    }
    void release() {
        if (rc && --*rc == 0) { std::free(owner); std::free(rc); }
        rc = 0; owner = 0; data = 0; rows = cols = 0; step = 0;
    }
    Mat clone() const {
        Mat m(rows, cols, flags);
        for (int y = 0; y < rows; y++) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * esz(flags));
        return m;
    }
    void copyTo(Mat& m) const { Mat c = clone(); m = c; }
    Mat operator()(const Rect& r) const {
        Mat m(*this);
        m.data = data + (size_t)r.y * step + (size_t)r.x * esz(flags);
        m.rows = r.height; m.cols = r.width;
        return m;
    }
    Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * esz(flags); m.cols = b - a; return m; }
    Mat row(int y) const { return rowRange(y, y + 1); }
    int type() const { return flags; }
    bool empty() const { return data == 0 || rows * cols == 0; }
    size_t elemSize() const { return esz(flags); }
    size_t step1() const { return step / esz(flags); }
    bool isContinuous() const { return step == (size_t)cols * esz(flags); }
    Size size() const { return Size(cols, rows); }
    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
    template <typename T> T& at(int y, int x) { return ((T*)(data + (size_t)y * step))[x]; }
    template <typename T> const T& at(int y, int x) const { return ((const T*)(data + (size_t)y * step))[x]; }
    static size_t esz(int type) { return (type == CV_32S || type == CV_32F) ? 4 : 1; }
private:
    int* rc;
    uchar* owner;
};

class _InputArray {
public:
    _InputArray() : m(0) {}
    _InputArray(const Mat& _m) : m(&_m) {}
    bool empty() const { return !m || m->empty(); }
    Mat getMat() const { return m ? *m : Mat(); }
protected:
    const Mat* m;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() : mm(0) {}
    _OutputArray(Mat& _m) : _InputArray(_m), mm(&_m) {}
    void create(int rows, int cols, int type) const { if (mm) mm->create(rows, cols, type); }
    void create(Size s, int type) const { create(s.height, s.width, type); }
    void release() const { if (mm) mm->release(); }
    Mat getMat() const { return mm ? *mm : Mat(); }
    Mat& getMatRef() const { return *mm; }
private:
    Mat* mm;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
inline InputArray noArray() { static _InputArray none; return none; }

float fastAtan2(float y, float x);
}  // namespace cv
#endif
