// Test driver that uses the drop-in ORB_SLAM2::ORBextractor exactly like Frame::ExtractORB does
// (S/Frame.cc:360-371): (*extractor)(im, cv::Mat(), mvKeys, mDescriptors).
// usage: shim_extract <in.raw> <w> <h> <out.bin>   (out: int32 n, n*28 keypoint bytes, n*32 descriptor bytes,
//                                                    then for each level: int32 w,h and the bordered level pixels)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ORBextractor.h"

int main(int argc, char** argv)
{
    if (argc < 5) return 2;
    const int w = std::atoi(argv[2]), h = std::atoi(argv[3]);
    std::vector<unsigned char> buf((size_t)w * h);
    FILE* f = std::fopen(argv[1], "rb");
    if (!f || std::fread(&buf[0], 1, buf.size(), f) != buf.size()) return 3;
    std::fclose(f);
    cv::Mat im(h, w, CV_8UC1, &buf[0]);

    ORB_SLAM2::ORBextractor* mpORBextractorLeft = new ORB_SLAM2::ORBextractor(1000, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> mvKeys;
    cv::Mat mDescriptors;
    (*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors);

    FILE* o = std::fopen(argv[4], "wb");
    int n = (int)mvKeys.size();
    std::fwrite(&n, 4, 1, o);
    if (n) std::fwrite(&mvKeys[0], sizeof(cv::KeyPoint), n, o);
    for (int i = 0; i < n; i++) std::fwrite(mDescriptors.ptr(i), 1, 32, o);
    for (int l = 0; l < mpORBextractorLeft->GetLevels(); l++) {
        const cv::Mat& m = mpORBextractorLeft->mvImagePyramid[l];
        int lw = m.cols, lh = m.rows;
        std::fwrite(&lw, 4, 1, o); std::fwrite(&lh, 4, 1, o);
        for (int y = -19; y < lh + 19; y++) std::fwrite(m.data + (ptrdiff_t)y * (ptrdiff_t)m.step - 19, 1, lw + 38, o);
    }
    std::fclose(o);
    std::printf("%d keypoints, scale[1]=%.7f\n", n, mpORBextractorLeft->GetScaleFactors()[1]);
    delete mpORBextractorLeft;
    return 0;
}
