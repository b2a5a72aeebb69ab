// ORBextractor.h -- drop-in replacement for ORB_SLAM2/include/ORBextractor.h of
// serviceberry3/weiner_slamit_v2: same namespace, class name, constructor, call operator, getters
// and public members (reference header lines 45-111), so Tracking.cc:156-162 and Frame.cc:360-371
// compile unchanged.  The body runs on a B200 through the C ABI in include/orb_b200.h; there is no
// CPU path.  Build notes: INTEGRATION.md.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <vector>
#include <opencv/cv.h>

struct orbb200_extractor;

namespace ORB_SLAM2
{

class ORBextractor
{
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();

    // ORB features + descriptors of one 8-bit grey image, dispersed by the quadtree.
    // The mask is ignored, as in the reference.
    void operator()(cv::InputArray image, cv::InputArray mask,
                    std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors);

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // Level images with their 19-px REFLECT_101 frame, refreshed on every call like the reference
    // (only Frame::ComputeStereoMatches reads them).  SetExportPyramid(false) skips the read-back.
    std::vector<cv::Mat> mvImagePyramid;
    void SetExportPyramid(bool on) { mbExportPyramid = on; }

    // B200 additions (not in the reference): CUDA device ordinal and Gaussian tap set
    // (0: OpenCV >= 3 taps, 1: OpenCV 2.4.9 taps) used when the device handle is (re)created.
    void SetDevice(int device) { mnDevice = device; }
    void SetBlurTaps(int taps) { mnBlurTaps = taps; }
    // the device handle of the last call (0 before the first one): Frame::ComputeStereoMatches reads the pyramids
    // where they lie in device memory through orbb200_extractor_pyramid_view
    orbb200_extractor* Handle() const { return mpHandle; }

protected:
    bool EnsureHandle(int width, int height);

    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;

    std::vector<int> mnFeaturesPerLevel;
    std::vector<int> umax;

    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

    orbb200_extractor* mpHandle;
    int mnHandleWidth, mnHandleHeight, mnDevice, mnBlurTaps;
    bool mbExportPyramid;

private:
    ORBextractor(const ORBextractor&);
    ORBextractor& operator=(const ORBextractor&);
};

} // namespace ORB_SLAM2

#endif
