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

// ---- the matrix expressions SearchByProjection(CurrentFrame, LastFrame, ...) evaluates: Rcw.t(), -expr,
// expr*Mat, Mat*Mat, expr+Mat (S/ORBmatcher.cc:1342-1363).  One MatOp that means
//     m = alpha * op(a) * b + c        (op = transpose when flags & 1; b, c optional)
// on small CV_32F matrices, with cv::gemm's small-matrix arithmetic: float products and sums in source order
// (tests/test_oracle_primitives.py pins that against cv2.gemm).
namespace cv {

MatOp::MatOp() {}
MatOp::~MatOp() {}

namespace {
class MiniOp : public MatOp {
public:
    void assign(const MatExpr& e, Mat& m, int = -1) const
    {
        const Mat& A = e.a;
        if (e.flags & 2) {                                   // elementwise a - b
            Mat d;
            d.create(A.rows, A.cols, CV_32F);
            for (int r = 0; r < A.rows; r++) for (int c = 0; c < A.cols; c++) d.at<float>(r, c) = A.at<float>(r, c) - e.b.at<float>(r, c);
            m = d;
            return;
        }
        const bool tr = (e.flags & 1) != 0;
        const int ar = tr ? A.cols : A.rows, ac = tr ? A.rows : A.cols;
        auto a_at = [&](int r, int c) { return tr ? A.at<float>(c, r) : A.at<float>(r, c); };
        const float alpha = (float)e.alpha;
        Mat out;
        if (e.b.empty()) {
            out.create(ar, ac, CV_32F);
            for (int r = 0; r < ar; r++) for (int c = 0; c < ac; c++) out.at<float>(r, c) = alpha * a_at(r, c);
        } else {
            const Mat& B = e.b;
            out.create(ar, B.cols, CV_32F);
            for (int r = 0; r < ar; r++)
                for (int c = 0; c < B.cols; c++) {
                    volatile float t = a_at(r, 0) * B.at<float>(0, c);
                    for (int k = 1; k < ac; k++) { volatile float p = a_at(r, k) * B.at<float>(k, c); t = t + p; }
                    volatile float v = alpha == 1.f ? (float)t : alpha * t;
                    if (!e.c.empty()) v = v + e.c.at<float>(r, c);
                    out.at<float>(r, c) = v;
                }
        }
        m = out;
    }
};
const MiniOp g_miniOp;
}  // namespace

MatExpr Mat::t() const { return MatExpr(&g_miniOp, 1, *this); }
MatExpr operator-(const MatExpr& e) { MatExpr r(e); r.alpha = -r.alpha; return r; }
MatExpr operator*(const MatExpr& e, const Mat& m) { MatExpr r(e); r.b = m; return r; }
MatExpr operator*(const Mat& a, const Mat& b) { return MatExpr(&g_miniOp, 0, a, b); }
MatExpr operator+(const MatExpr& e, const Mat& m) { MatExpr r(e); r.c = m; return r; }
MatExpr operator-(const Mat& a, const Mat& b) { return MatExpr(&g_miniOp, 2, a, b); }
// scalings: evaluated as a * (float)alpha, which is what convertTo does for CV_32F (cvtScale with float work type)
MatExpr operator-(const Mat& a) { return MatExpr(&g_miniOp, 0, a, Mat(), Mat(), -1.0); }
MatExpr operator/(const Mat& a, double s) { return MatExpr(&g_miniOp, 0, a, Mat(), Mat(), 1. / s); }
MatExpr operator*(double s, const Mat& a) { return MatExpr(&g_miniOp, 0, a, Mat(), Mat(), s); }
MatExpr operator*(double s, const MatExpr& e) { MatExpr r(e); r.alpha *= s; return r; }

// ---- what Frame::ComputeStereoMatches evaluates on its 11 x 11 patches (S/Frame.cc:684-707): convertTo(CV_32F) of a
// CV_8U view (in place), c * Mat::ones(...), Mat - expr, norm(a, b, NORM_L1).  Every value is a small integer held in
// a float, so each of these is exact whatever the evaluation order.
MatExpr Mat::ones(int rows, int cols, int type)
{
    CV_Assert(type == CV_32F);
    Mat m;
    m.create(rows, cols, CV_32F);
    for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) m.at<float>(r, c) = 1.f;
    return MatExpr(&g_miniOp, 0, m);
}
MatExpr operator-(const Mat& a, const MatExpr& e)
{
    Mat b;
    e.op->assign(e, b);
    return MatExpr(&g_miniOp, 2, a, b);
}
void Mat::convertTo(OutputArray dst, int rtype, double alpha, double beta) const
{
    CV_Assert(type() == CV_8U && rtype == CV_32F && alpha == 1 && beta == 0);
    const Mat src = *this;                         // dst may be *this (IL.convertTo(IL, CV_32F))
    Mat out;
    out.create(src.rows, src.cols, CV_32F);
    for (int r = 0; r < src.rows; r++) for (int c = 0; c < src.cols; c++) out.at<float>(r, c) = (float)src.at<uchar>(r, c);
    *static_cast<Mat*>(dst.getObj()) = out;
}
double norm(InputArray src1, InputArray src2, int normType, InputArray)
{
    const Mat& a = *static_cast<const Mat*>(src1.getObj());
    const Mat& b = *static_cast<const Mat*>(src2.getObj());
    CV_Assert(normType == NORM_L1 && a.type() == CV_32F && b.type() == CV_32F && a.rows == b.rows && a.cols == b.cols);
    double s = 0;
    for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) s += std::fabs((double)a.at<float>(r, c) - (double)b.at<float>(r, c));
    return s;
}

// Mat::dot for continuous CV_32F: products and sum in double, element order (core/src/matmul.cpp, dotProd_ for short vectors)
double Mat::dot(InputArray other) const
{
    const Mat& o = *static_cast<const Mat*>(other.getObj());
    CV_Assert(type() == CV_32F && o.type() == CV_32F && rows == o.rows && cols == o.cols);
    double r = 0;
    for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) r += (double)at<float>(y, x) * (double)o.at<float>(y, x);
    return r;
}

InputOutputArray noArray() { static _InputOutputArray none; return none; }

// cv::norm, NORM_L2 of a continuous CV_32F array: squares accumulated in double in element order (core/src/stat.cpp, normL2_)
double norm(InputArray src, int normType, InputArray)
{
    const Mat& m = *static_cast<const Mat*>(src.getObj());
    CV_Assert(normType == NORM_L2 && m.type() == CV_32F && m.isContinuous());
    double s = 0;
    const float* p = m.ptr<float>();
    for (size_t i = 0; i < m.total(); i++) s += (double)p[i] * (double)p[i];
    return std::sqrt(s);
}

}  // namespace cv
