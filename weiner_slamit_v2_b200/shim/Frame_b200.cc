// Frame_b200.cc -- B200 bodies for the two Frame methods of the "next" row N1:
//   Frame::UndistortKeyPoints   (replaces S/Frame.cc:529-559)
//   Frame::ComputeImageBounds   (replaces S/Frame.cc:561-589)
// Compiles against the reference's own Frame.h; guard the two reference bodies with
// #ifndef ORB_B200_FRAME and add this file.  cv::undistortPoints runs on the device.
#include "Frame.h"

#include <cstdio>
#include <vector>

#include "orb_b200.h"

namespace ORB_SLAM2
{

namespace
{
struct ThreadHandle {
    orbb200_matcher* h;
    ThreadHandle() : h(0) {}
    ~ThreadHandle() { if (h) orbb200_matcher_destroy(h); }
    orbb200_matcher* get()
    {
        if (!h && orbb200_matcher_create(1, 4096, 0, &h) != ORBB200_OK) {
            std::fprintf(stderr, "Frame(B200): %s\n", orbb200_last_error());
            h = 0;
        }
        return h;
    }
};
thread_local ThreadHandle tlsHandle;

void CameraModel(const cv::Mat& K, const cv::Mat& D, float k[4], float d[5])
{
    k[0] = K.at<float>(0, 0); k[1] = K.at<float>(1, 1); k[2] = K.at<float>(0, 2); k[3] = K.at<float>(1, 2);
    for (int i = 0; i < 5; i++) d[i] = i < (int)D.total() ? D.at<float>(i) : 0.f;
}
}  // namespace

void Frame::UndistortKeyPoints()
{
    if (mDistCoef.at<float>(0) == 0.0) {
        mvKeysUn = mvKeys;
        return;
    }
    std::vector<float> xy(2 * (size_t)N), out(2 * (size_t)N);
    for (int i = 0; i < N; i++) { xy[2 * i] = mvKeys[i].pt.x; xy[2 * i + 1] = mvKeys[i].pt.y; }
    float k[4], d[5];
    CameraModel(mK, mDistCoef, k, d);
    orbb200_matcher* h = tlsHandle.get();
    if (!h || (N > 0 && orbb200_undistort_points(h, &xy[0], &out[0], N, k, d) != ORBB200_OK)) {
        std::fprintf(stderr, "Frame(B200)::UndistortKeyPoints: %s\n", orbb200_last_error());
        out = xy;
    }
    mvKeysUn.resize(N);
    for (int i = 0; i < N; i++) {
        cv::KeyPoint kp = mvKeys[i];
        kp.pt.x = out[2 * i];
        kp.pt.y = out[2 * i + 1];
        mvKeysUn[i] = kp;
    }
}

void Frame::ComputeImageBounds(const cv::Mat& imLeft)
{
    float k[4], d[5], b[4] = {0.f, 0.f, (float)imLeft.cols, (float)imLeft.rows};
    CameraModel(mK, mDistCoef, k, d);
    orbb200_matcher* h = tlsHandle.get();
    if (!h || orbb200_image_bounds(h, imLeft.cols, imLeft.rows, k, d, b) != ORBB200_OK)
        std::fprintf(stderr, "Frame(B200)::ComputeImageBounds: %s\n", orbb200_last_error());
    mnMinX = b[0]; mnMinY = b[1]; mnMaxX = b[2]; mnMaxY = b[3];
}

}  // namespace ORB_SLAM2
