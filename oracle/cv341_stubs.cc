// cv341_stubs.cc -- TEST INFRASTRUCTURE.  The handful of out-of-line cv::Mat functions that the
// reference's ORBmatcher.cc / Frame.cc / MapPoint.cc actually EXECUTE on the three in-scope matcher paths
// (Mat::row, Mat::clone), written from scratch against the OpenCV 3.4.1 headers that ship in
// /root/reference/openCVLibrary341 (the matching libopencv_core is absent: .MISSING_LARGE_BLOBS).
// Everything else those objects reference is never reached and gets a dummy definition from
// gen_link_stubs.py.  2-D matrices only.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <opencv2/core/core.hpp>

namespace cv {

// sub-matrix header: share the data, narrow rows / columns
Mat::Mat(const Mat& m, const Range& rowRange, const Range& colRange)
    : flags(MAGIC_VAL), dims(0), rows(0), cols(0), data(0), datastart(0), dataend(0), datalimit(0), allocator(0), u(0), size(&rows)
{
    *this = m;
    if (rowRange != Range::all()) {
        data += step.p[0] * (size_t)rowRange.start;
        rows = rowRange.end - rowRange.start;
        flags |= SUBMATRIX_FLAG;
    }
    if (colRange != Range::all()) {
        data += elemSize() * (size_t)colRange.start;
        cols = colRange.end - colRange.start;
        flags |= SUBMATRIX_FLAG;
    }
    if (rows == 1 || step.p[0] == (size_t)cols * elemSize()) flags |= CONTINUOUS_FLAG;
    else flags &= ~CONTINUOUS_FLAG;
}

void Mat::create(int d, const int* sizes, int type)
{
    if (d != 2) { std::fprintf(stderr, "cv341_stubs: only 2-D matrices\n"); std::abort(); }
    type &= TYPE_MASK;
    if (data && dims == 2 && rows == sizes[0] && cols == sizes[1] && this->type() == type) return;
    release();
    flags = MAGIC_VAL | type | CONTINUOUS_FLAG;
    dims = 2; rows = sizes[0]; cols = sizes[1];
    const size_t esz = CV_ELEM_SIZE(type);
    step.p[0] = esz * (size_t)cols; step.p[1] = esz;
    const size_t total = step.p[0] * (size_t)rows;
    u = (UMatData*)std::calloc(1, sizeof(UMatData));        // plain storage: refcount + data pointers are all we use
    u->refcount = 1;
    u->origdata = u->data = (uchar*)std::malloc(total ? total : 1);
    u->size = total;
    data = u->data; datastart = u->data;
    datalimit = dataend = data + total;
}

void Mat::deallocate()
{
    if (u) { std::free(u->origdata); std::free(u); u = 0; }
}

void Mat::copyTo(OutputArray dst) const
{
    Mat& d = *(Mat*)dst.getObj();                             // Mat::clone() is the only caller here
    if (empty()) { d.release(); return; }
    d.create(rows, cols, type());
    const size_t rowBytes = (size_t)cols * elemSize();
    for (int y = 0; y < rows; y++) std::memcpy(d.data + d.step.p[0] * (size_t)y, data + step.p[0] * (size_t)y, rowBytes);
}

}  // namespace cv
