// mini-cv (test infrastructure): features2d declarations used by ORBextractor.cc.
#ifndef MINICV_FEATURES2D_HPP
#define MINICV_FEATURES2D_HPP
#include <opencv2/core/core.hpp>
namespace cv {
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
class KeyPointsFilter {
public:
    static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);
};
}
#endif
