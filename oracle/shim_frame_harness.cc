// shim_frame_harness.cc -- TEST INFRASTRUCTURE, part of oracle/_ref/libshim_matcher.so only.
// Drives the drop-in classes the way the reference's own code drives them: the reference's UNMODIFIED Frame
// constructors (S/Frame.cc:70-133 stereo, :135-188 RGB-D) are called with ORB_SLAM2::ORBextractor objects whose
// body is weiner_slamit_v2_b200/shim/ORBextractor.cc, so Frame::ExtractORB (S/Frame.cc:360-371), the two
// std::threads of the stereo constructor (:93-96), UndistortKeyPoints, ComputeStereoMatches, ComputeImageBounds
// and AssignFeaturesToGrid run in the reference's order on real objects, and ORBmatcher (shim body) is then
// called on those Frames exactly as S/Tracking.cc:799-800 calls it.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>
#include <opencv2/core/core.hpp>

#include <map>
#include <string>

#include "Frame.h"
#include "ORBVocabulary.h"
#include "ORBmatcher.h"

using namespace ORB_SLAM2;

namespace {
cv::Mat camera_matrix(const float* K4)
{
    cv::Mat K(3, 3, CV_32F);
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) K.at<float>(r, c) = r == c ? 1.f : 0.f;
    K.at<float>(0, 0) = K4[0]; K.at<float>(1, 1) = K4[1]; K.at<float>(0, 2) = K4[2]; K.at<float>(1, 2) = K4[3];
    return K;
}
cv::Mat dist_coef(const float* d5)
{
    cv::Mat D(5, 1, CV_32F);
    for (int i = 0; i < 5; i++) D.at<float>(i) = d5 ? d5[i] : 0.f;
    return D;
}
void copy_frame(const Frame& F, int cap, int32_t* n, void* keys28, void* keysUn28, uint8_t* desc)
{
    *n = F.N;
    const int m = F.N < cap ? F.N : cap;
    static_assert(sizeof(cv::KeyPoint) == 28, "cv::KeyPoint layout");
    if (m > 0) {
        std::memcpy(keys28, &F.mvKeys[0], 28 * (size_t)m);
        std::memcpy(keysUn28, &F.mvKeysUn[0], 28 * (size_t)m);
        for (int i = 0; i < m; i++) std::memcpy(desc + 32 * (size_t)i, F.mDescriptors.ptr(i), 32);
    }
}
}  // namespace

extern "C" {

// Two frames through the RGB-D constructor (no Posenet on that one), then ORBmatcher(ratio, true).SearchForInitialization
// with vbPrevMatched = F1's undistorted keypoints, as Tracking::MonocularInitialization does (S/Tracking.cc:773-800).
int shimm_rgbd_frames_init_match(const uint8_t* imgA, const uint8_t* imgB, int w, int h, int nfeatures,
                                 const float* K4, const float* dist5, float ratio, int window, int cap,
                                 int32_t* nA, void* keysA, void* keysUnA, uint8_t* descA,
                                 int32_t* nB, void* keysB, void* keysUnB, uint8_t* descB,
                                 float* bounds4, int32_t* matches12, float* prev_matched)
{
    ORBextractor ex(nfeatures, 1.2f, 8, 20, 7);
    ex.SetExportPyramid(false);
    cv::Mat K = camera_matrix(K4), D = dist_coef(dist5);
    cv::Mat a(h, w, CV_8UC1, (void*)imgA), b(h, w, CV_8UC1, (void*)imgB);
    cv::Mat depth(h, w, CV_32F);
    std::memset(depth.data, 0, (size_t)w * h * 4);
    Frame::mbInitialComputations = true;
    Frame F1(a, depth, 0.0, &ex, static_cast<ORBVocabulary*>(NULL), K, D, 40.f, 35.f);
    Frame F2(b, depth, 1.0, &ex, static_cast<ORBVocabulary*>(NULL), K, D, 40.f, 35.f);
    copy_frame(F1, cap, nA, keysA, keysUnA, descA);
    copy_frame(F2, cap, nB, keysB, keysUnB, descB);
    bounds4[0] = Frame::mnMinX; bounds4[1] = Frame::mnMinY; bounds4[2] = Frame::mnMaxX; bounds4[3] = Frame::mnMaxY;
    if (F1.N > cap || F2.N > cap) return -1;
    std::vector<cv::Point2f> prev(F1.mvKeysUn.size());
    for (size_t i = 0; i < prev.size(); i++) prev[i] = F1.mvKeysUn[i].pt;
    std::vector<int> m12;
    ORBmatcher matcher(ratio, true);
    const int n = matcher.SearchForInitialization(F1, F2, prev, m12, window);
    for (int i = 0; i < F1.N; i++) { matches12[i] = m12[i]; prev_matched[2 * i] = prev[i].x; prev_matched[2 * i + 1] = prev[i].y; }
    return n;
}

// One stereo Frame through the reference's stereo constructor: two extractors on two std::threads, then
// Frame::ComputeStereoMatches on the pyramids the extractors left on the device.
int shimm_stereo_frame(const uint8_t* left, const uint8_t* right, int w, int h, int nfeatures, const float* K4,
                       float bf, int cap, int32_t* nL, void* keysL, uint8_t* descL, int32_t* nR, void* keysR, uint8_t* descR,
                       float* u_right, float* depth)
{
    ORBextractor exL(nfeatures, 1.2f, 8, 20, 7), exR(nfeatures, 1.2f, 8, 20, 7);
    exL.SetExportPyramid(false); exR.SetExportPyramid(false);
    cv::Mat K = camera_matrix(K4), D = dist_coef(NULL);
    cv::Mat a(h, w, CV_8UC1, (void*)left), b(h, w, CV_8UC1, (void*)right);
    Frame::mbInitialComputations = true;
    // The constructor reads mb (minZ = mb, S/Frame.cc:620) before it assigns it (:130): build the Frame in zeroed storage
    // so that the value it reads is 0 (maxD = +inf) and the run is reproducible.
    void* mem = std::calloc(1, sizeof(Frame));
    Frame& F = *new (mem) Frame(a, b, 0.0, &exL, &exR, static_cast<ORBVocabulary*>(NULL), K, D, bf, 35.f);
    struct Cleanup { Frame* f; void* m; ~Cleanup() { f->~Frame(); std::free(m); } } cleanup = {&F, mem};
    *nL = F.N; *nR = (int32_t)F.mvKeysRight.size();
    if (F.N > cap || *nR > cap) return -1;
    int cnt = 0;
    for (int i = 0; i < F.N; i++) {
        std::memcpy((char*)keysL + 28 * (size_t)i, &F.mvKeys[i], 28);
        std::memcpy(descL + 32 * (size_t)i, F.mDescriptors.ptr(i), 32);
        u_right[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i];
        cnt += F.mvuRight[i] != -1.0f;
    }
    for (int i = 0; i < *nR; i++) {
        std::memcpy((char*)keysR + 28 * (size_t)i, &F.mvKeysRight[i], 28);
        std::memcpy(descR + 32 * (size_t)i, F.mDescriptorsRight.ptr(i), 32);
    }
    return cnt;
}

// Tracking's extractor sequence (S/Tracking.cc:156-162, 268-271; Reset() returns to the first): the 2*nFeatures
// initialisation extractor, then the nFeatures one, then the first again, all on one thread.  Returns the three counts.
int shimm_extractor_sequence(const uint8_t* img, int w, int h, int nfeatures, int32_t* counts3)
{
    ORBextractor ini(2 * nfeatures, 1.2f, 8, 20, 7), left(nfeatures, 1.2f, 8, 20, 7);
    ini.SetExportPyramid(false); left.SetExportPyramid(false);
    cv::Mat a(h, w, CV_8UC1, (void*)img), d;
    std::vector<cv::KeyPoint> k;
    ini(a, cv::Mat(), k, d); counts3[0] = (int32_t)k.size();
    left(a, cv::Mat(), k, d); counts3[1] = (int32_t)k.size();
    ini(a, cv::Mat(), k, d); counts3[2] = (int32_t)k.size();
    return 0;
}

// Frame::ComputeBoW (shim body: the DBoW2 transform on the device) on a Frame that holds `n` descriptors and a
// vocabulary read by the reference's own loadFromTextFile.  Outputs flattened in std::map order.  -1: file did not load.
int shimm_frame_compute_bow(const char* path, int n, const uint8_t* desc,
                            int32_t* bow_n, uint32_t* bow_word, double* bow_value,
                            int32_t* fv_n, uint32_t* fv_node, int32_t* fv_start, uint32_t* fv_feat)
{
    static std::map<std::string, ORBVocabulary*> cache;
    ORBVocabulary*& voc = cache[path];
    if (!voc) {
        voc = new ORBVocabulary();
        if (!voc->loadFromTextFile(path)) { delete voc; voc = NULL; cache.erase(path); return -1; }
    }
    Frame F;
    F.mpORBvocabulary = voc;
    F.mDescriptors = cv::Mat(n > 0 ? n : 1, 32, CV_8U, (void*)desc).rowRange(0, n);
    F.ComputeBoW();
    int k = 0;
    for (DBoW2::BowVector::const_iterator it = F.mBowVec.begin(); it != F.mBowVec.end(); ++it, ++k) { bow_word[k] = it->first; bow_value[k] = it->second; }
    *bow_n = k;
    int a = 0, pos = 0;
    for (DBoW2::FeatureVector::const_iterator it = F.mFeatVec.begin(); it != F.mFeatVec.end(); ++it, ++a) {
        fv_node[a] = it->first; fv_start[a] = pos;
        for (size_t j = 0; j < it->second.size(); j++) fv_feat[pos++] = it->second[j];
    }
    fv_start[a] = pos;
    *fv_n = a;
    return 0;
}

}  // extern "C"
