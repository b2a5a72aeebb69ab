// cv341_stubs_shim.cc -- TEST INFRASTRUCTURE, part of oracle/_ref/libshim_matcher.so only.
// The out-of-line OpenCV 3.4.1 functions that weiner_slamit_v2_b200/shim/ORBextractor.cc and the reference's
// Frame constructors execute when they are driven by shim_frame_harness.cc: the cv::InputArray / cv::OutputArray
// proxies over a cv::Mat and the Rect sub-matrix header.  Written from scratch against the 3.4.1 headers under
// /root/reference/openCVLibrary341 (the matching libopencv_core is absent: .MISSING_LARGE_BLOBS).  cv::Mat only.
#include <cstdio>
#include <cstdlib>
#include <opencv2/core/core.hpp>

namespace cv {

static void only_mat(int flags, const char* who)
{
    const int k = flags & _InputArray::KIND_MASK;
    if (k != _InputArray::MAT && k != _InputArray::NONE) { std::fprintf(stderr, "cv341_stubs_shim: %s on a non-Mat array\n", who); std::abort(); }
}

int _InputArray::kind() const { return flags & KIND_MASK; }

Mat _InputArray::getMat_(int) const
{
    only_mat(flags, "getMat_");
    return (flags & KIND_MASK) == MAT ? *static_cast<const Mat*>(obj) : Mat();
}

bool _InputArray::empty() const
{
    only_mat(flags, "empty");
    return (flags & KIND_MASK) == NONE || static_cast<const Mat*>(obj)->empty();
}

void _OutputArray::create(int rows, int cols, int mtype, int, bool, int) const
{
    only_mat(flags, "create");
    static_cast<Mat*>(obj)->create(rows, cols, mtype);
}

void _OutputArray::release() const
{
    only_mat(flags, "release");
    if ((flags & KIND_MASK) == MAT) static_cast<Mat*>(obj)->release();
}

Mat::Mat(const Mat& m, const Rect& roi)
    : flags(m.flags), dims(2), rows(roi.height), cols(roi.width), data(m.data + roi.y * m.step.p[0]),
      datastart(m.datastart), dataend(m.dataend), datalimit(m.datalimit), allocator(m.allocator), u(m.u), size(&rows)
{
    const size_t esz = CV_ELEM_SIZE(flags);
    data += roi.x * esz;
    if (u) CV_XADD(&u->refcount, 1);
    if (roi.width < m.cols || roi.height < m.rows) flags |= SUBMATRIX_FLAG;
    step.p[0] = m.step.p[0]; step.p[1] = esz;
    if (rows == 1 || step.p[0] == (size_t)cols * esz) flags |= CONTINUOUS_FLAG;
    else flags &= ~CONTINUOUS_FLAG;
}

}  // namespace cv
