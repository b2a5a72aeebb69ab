// ORBextractor.cc -- ORB_SLAM2::ORBextractor as a thin shim over liborb_b200.so.
// Replaces ORB_SLAM2/src/ORBextractor.cc; every pixel operation runs in CUDA (sm_100a).
// Error convention: the reference never throws on this path, so a device failure
// is logged to stderr and the call returns with empty outputs.
#include "ORBextractor.h"

#include <cassert>
#include <cstdio>
#include <cstring>

#include "orb_b200.h"

namespace ORB_SLAM2
{

static const int kEdge = 19;   // EDGE_THRESHOLD: width of the frame kept around every pyramid level

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST),
      minThFAST(_minThFAST), mpHandle(0), mnHandleWidth(0), mnHandleHeight(0), mnDevice(0), mnBlurTaps(0),
      mbExportPyramid(true)
{
    // The scale tables are pure functions of the five parameters; take them from the library (which
    // evaluates them with the reference's float/double mix) via a throw-away handle geometry the
    // first time an image arrives.  Until then size the vectors so the getters are well formed.
    mvScaleFactor.assign(nlevels, 1.0f);
    mvInvScaleFactor.assign(nlevels, 1.0f);
    mvLevelSigma2.assign(nlevels, 1.0f);
    mvInvLevelSigma2.assign(nlevels, 1.0f);
    mnFeaturesPerLevel.assign(nlevels, 0);
    umax.assign(16, 0);
    mvImagePyramid.resize(nlevels);
    for (int i = 1; i < nlevels; i++) {                 // same recurrence as the library (float * double)
        mvScaleFactor[i] = (float)(mvScaleFactor[i - 1] * scaleFactor);
        mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
    }
    for (int i = 0; i < nlevels; i++) {
        mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i];
        mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i];
    }
}

ORBextractor::~ORBextractor()
{
    if (mpHandle) orbb200_extractor_destroy(mpHandle);
}

bool ORBextractor::EnsureHandle(int width, int height)
{
    if (mpHandle && width == mnHandleWidth && height == mnHandleHeight) return true;
    if (mpHandle) { orbb200_extractor_destroy(mpHandle); mpHandle = 0; }
    int rc = orbb200_extractor_create(nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST, width, height,
                                      1, mnDevice, mnBlurTaps, &mpHandle);
    if (rc != ORBB200_OK) {
        std::fprintf(stderr, "ORBextractor(B200): cannot create device handle: %s\n", orbb200_last_error());
        mpHandle = 0;
        return false;
    }
    mnHandleWidth = width; mnHandleHeight = height;
    orbb200_extractor_tables(mpHandle, &mvScaleFactor[0], &mvInvScaleFactor[0], &mvLevelSigma2[0],
                             &mvInvLevelSigma2[0], &mnFeaturesPerLevel[0], &umax[0]);
    return true;
}

static inline int Reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
}

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*mask*/,
                              std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors)
{
    if (_image.empty()) return;                       // outputs untouched, like the reference

    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1);

    _keypoints.clear();
    if (!EnsureHandle(image.cols, image.rows)) { _descriptors.release(); return; }

    // cv::KeyPoint and orbb200_keypoint share one 28-byte layout: let the library fill the vector.
    static_assert(sizeof(cv::KeyPoint) == sizeof(orbb200_keypoint), "cv::KeyPoint layout changed");
    const int cap = orbb200_extractor_max_keypoints(mpHandle);
    std::vector<cv::KeyPoint> keys(cap);
    std::vector<unsigned char> desc((size_t)cap * 32);
    int32_t n = 0;
    int rc = orbb200_extract_host(mpHandle, image.data, 1, image.step, image.step * (size_t)image.rows,
                                  reinterpret_cast<orbb200_keypoint*>(&keys[0]), &desc[0], &n, cap);
    if (rc != ORBB200_OK) {
        std::fprintf(stderr, "ORBextractor(B200): extraction failed: %s\n", orbb200_last_error());
        _descriptors.release();
        return;
    }

    if (n == 0) {
        _descriptors.release();
    } else {
        _descriptors.create(n, 32, CV_8U);
        cv::Mat out = _descriptors.getMat();
        for (int i = 0; i < n; i++) std::memcpy(out.ptr(i), &desc[(size_t)i * 32], 32);
    }
    keys.resize(n);
    _keypoints.swap(keys);

    if (mbExportPyramid) {
        for (int level = 0; level < nlevels; ++level) {
            int w = 0, h = 0;
            orbb200_extractor_level_size(mpHandle, level, &w, &h);
            cv::Mat whole(h + 2 * kEdge, w + 2 * kEdge, CV_8UC1);
            cv::Mat inner = whole(cv::Rect(kEdge, kEdge, w, h));
            if (orbb200_extractor_get_level(mpHandle, 0, level, 0, inner.data, inner.step) != ORBB200_OK) break;
            for (int y = -kEdge; y < h + kEdge; y++) {      // REFLECT_101 frame around the level
                unsigned char* row = inner.data + (ptrdiff_t)y * (ptrdiff_t)inner.step;
                const unsigned char* src = inner.data + (ptrdiff_t)Reflect101(y, h) * (ptrdiff_t)inner.step;
                if (y < 0 || y >= h) std::memcpy(row, src, w);
                for (int x = 1; x <= kEdge; x++) { row[-x] = src[Reflect101(-x, w)]; row[w - 1 + x] = src[Reflect101(w - 1 + x, w)]; }
            }
            mvImagePyramid[level] = inner;
        }
    }
}

} // namespace ORB_SLAM2
