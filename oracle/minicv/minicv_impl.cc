// mini-cv implementation (TEST INFRASTRUCTURE): the five OpenCV image functions that
// ORBextractor.cc calls, implemented on the C oracle primitives (each pinned against cv2 4.13).
#include <opencv/cv.h>
#include <algorithm>
#include "orb_oracle.h"

namespace cv {

static int g_blur_variant = 0;
extern "C" void minicv_set_blur_variant(int v) { g_blur_variant = v; }

float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression)
{
    Mat img = image.getMat();
    keypoints.clear();
    if (img.empty()) return;
    int cap = nonmaxSuppression ? (img.cols / 2 + 2) * (img.rows / 2 + 2) : img.cols * img.rows;
    std::vector<orc_keypoint> tmp(cap > 0 ? cap : 1);
    int n = orc_fast9_16(img.data, img.cols, img.rows, (int)img.step, threshold, nonmaxSuppression ? 1 : 0, tmp.data(), cap);
    keypoints.reserve(n);
    for (int i = 0; i < n; i++)
        keypoints.push_back(KeyPoint(tmp[i].x, tmp[i].y, tmp[i].size, tmp[i].angle, tmp[i].response, tmp[i].octave, tmp[i].class_id));
}

void KeyPointsFilter::retainBest(std::vector<KeyPoint>& keypoints, int npoints)
{
    // only reached from the reference's dead ComputeKeyPointsOld()
    if (npoints >= 0 && (int)keypoints.size() > npoints) {
        std::stable_sort(keypoints.begin(), keypoints.end(),
                         [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
        keypoints.resize(npoints);
    }
}

void resize(InputArray _src, OutputArray _dst, Size dsize, double, double, int)
{
    Mat src = _src.getMat();
    _dst.create(dsize.height, dsize.width, src.type());
    Mat dst = _dst.getMat();
    orc_resize_linear_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right, int)
{
    Mat src = _src.getMat();
    // the source may be the interior of the destination (ORBextractor.cc:1159): work from a copy
    Mat tmp = src.clone();
    _dst.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = _dst.getMat();
    (void)bottom; (void)right; // the reference always passes the same width on all four sides
    orc_copy_make_border_reflect101(tmp.data, tmp.cols, tmp.rows, (int)tmp.step, dst.data, (int)dst.step, top);
}

void GaussianBlur(InputArray _src, OutputArray _dst, Size, double, double, int)
{
    Mat src = _src.getMat();
    Mat tmp = src.clone();
    _dst.create(src.rows, src.cols, src.type());
    Mat dst = _dst.getMat();
    orc_gaussian_blur7(tmp.data, tmp.cols, tmp.rows, (int)tmp.step, dst.data, (int)dst.step, g_blur_variant);
}

}  // namespace cv
