// Frame_b200.cc -- B200 bodies for four Frame methods:
//   Frame::UndistortKeyPoints   (replaces S/Frame.cc:529-559)
//   Frame::ComputeImageBounds   (replaces S/Frame.cc:561-589)
//   Frame::ComputeBoW           (replaces S/Frame.cc:520-527: the DBoW2 transform runs on the device)
//   Frame::ComputeStereoMatches (replaces S/Frame.cc:591-763)
// Compiles against the reference's own Frame.h; guard the two reference bodies with
// #ifndef ORB_B200_FRAME and add this file.  cv::undistortPoints runs on the device.
#include "Frame.h"

#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
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

namespace
{
// The vocabulary tree lives in protected members of TemplatedVocabulary; a derived class may name them, and the
// pointers to members it forms are plain `T TemplatedVocabulary::*` values (no object of this type is ever created).
struct VocabularyFields : public ORBVocabulary {
    typedef std::vector<Node> Nodes;
    static Nodes ORBVocabulary::* NodesMember() { return &VocabularyFields::m_nodes; }
    static int ORBVocabulary::* LevelsMember() { return &VocabularyFields::m_L; }
    static orbb200_vocabulary* Upload(const ORBVocabulary* voc)
    {
        const Nodes& nodes = voc->*NodesMember();
        const int n = (int)nodes.size();
        std::vector<int32_t> childStart(n + 1, 0), children, wordId(n, -1);
        std::vector<unsigned char> desc((size_t)n * 32, 0);
        std::vector<double> weight(n, 0.0);
        for (int i = 0; i < n; i++) {
            childStart[i] = (int32_t)children.size();
            for (size_t c = 0; c < nodes[i].children.size(); c++) children.push_back((int32_t)nodes[i].children[c]);
            if (!nodes[i].descriptor.empty()) std::memcpy(&desc[(size_t)i * 32], nodes[i].descriptor.ptr<unsigned char>(), 32);
            if (nodes[i].children.empty()) { wordId[i] = (int32_t)nodes[i].word_id; weight[i] = nodes[i].weight; }
        }
        childStart[n] = (int32_t)children.size();
        if (children.empty()) children.push_back(0);
        orbb200_vocabulary* h = 0;
        if (orbb200_vocabulary_create(0, n, voc->*LevelsMember(), &childStart[0], &children[0], &desc[0], &wordId[0], &weight[0], &h) != ORBB200_OK) {
            std::fprintf(stderr, "Frame(B200): vocabulary upload: %s\n", orbb200_last_error());
            return 0;
        }
        return h;
    }
};

// one device copy per vocabulary object, made on first use (System loads the vocabulary once, S/System.cc:70-85)
orbb200_vocabulary* DeviceVocabulary(const ORBVocabulary* voc)
{
    static std::mutex mtx;
    static std::map<const ORBVocabulary*, orbb200_vocabulary*> cache;
    std::unique_lock<std::mutex> lock(mtx);
    std::map<const ORBVocabulary*, orbb200_vocabulary*>::iterator it = cache.find(voc);
    if (it != cache.end()) return it->second;
    return cache[voc] = VocabularyFields::Upload(voc);
}
}  // namespace

void Frame::ComputeBoW()
{
    if (!mBowVec.empty()) return;
    orbb200_matcher* h = tlsHandle.get();
    orbb200_vocabulary* voc = mpORBvocabulary ? DeviceVocabulary(mpORBvocabulary) : 0;
    const int n = mDescriptors.rows;
    if (!h || !voc || n == 0) return;
    std::vector<unsigned char> desc((size_t)n * 32);
    for (int j = 0; j < n; j++) std::memcpy(&desc[(size_t)j * 32], mDescriptors.ptr<unsigned char>(j), 32);   // Converter::toDescriptorVector
    int32_t nn = n, bowN = 0, fvN = 0;
    std::vector<uint32_t> word(n), node(n), feat(n);
    std::vector<double> value(n);
    std::vector<int32_t> start(n + 1);
    if (orbb200_bow_transform(h, voc, 1, &nn, &desc[0], n, 4, &bowN, &word[0], &value[0], &fvN, &node[0], &start[0], &feat[0], 0) != ORBB200_OK) {
        std::fprintf(stderr, "Frame(B200)::ComputeBoW: %s\n", orbb200_last_error());
        return;
    }
    for (int k = 0; k < bowN; k++) mBowVec.insert(mBowVec.end(), DBoW2::BowVector::value_type(word[k], value[k]));
    for (int a = 0; a < fvN; a++)
        for (int p = start[a]; p < start[a + 1]; p++) mFeatVec.addFeature(node[a], feat[p]);
}

void Frame::ComputeStereoMatches()
{
    mvuRight = std::vector<float>(N, -1.0f);
    mvDepth = std::vector<float>(N, -1.0f);
    const int Nr = (int)mvKeysRight.size();
    orbb200_matcher* h = tlsHandle.get();
    if (!h || N == 0 || Nr == 0) return;

    // mvKeys / mvKeysRight as structure-of-arrays; the descriptor matrices are continuous N x 32
    std::vector<float> lx(N), ly(N), rx(Nr), ry(Nr);
    std::vector<int32_t> lo(N), ro(Nr);
    for (int i = 0; i < N; i++) { lx[i] = mvKeys[i].pt.x; ly[i] = mvKeys[i].pt.y; lo[i] = mvKeys[i].octave; }
    for (int i = 0; i < Nr; i++) { rx[i] = mvKeysRight[i].pt.x; ry[i] = mvKeysRight[i].pt.y; ro[i] = mvKeysRight[i].octave; }
    const cv::Mat dl = mDescriptors.isContinuous() ? mDescriptors : mDescriptors.clone();
    const cv::Mat dr = mDescriptorsRight.isContinuous() ? mDescriptorsRight : mDescriptorsRight.clone();
    int32_t nl = N, nr = Nr, nmatches = 0;
    orbb200_frame_view left = {&nl, &lx[0], &ly[0], &lo[0], 0, dl.ptr<unsigned char>(), N};
    orbb200_frame_view right = {&nr, &rx[0], &ry[0], &ro[0], 0, dr.ptr<unsigned char>(), Nr};

    // the pyramids: where the two extractors left them on the device, else the host copies in mvImagePyramid
    orbb200_pyramid_view lp, rp;
    int where = ORBB200_DEVICE_PYRAMIDS;
    if (!mpORBextractorLeft->Handle() || !mpORBextractorRight->Handle() ||
        orbb200_extractor_pyramid_view(mpORBextractorLeft->Handle(), &lp) != ORBB200_OK ||
        orbb200_extractor_pyramid_view(mpORBextractorRight->Handle(), &rp) != ORBB200_OK) {
        where = 0;
        std::memset(&lp, 0, sizeof(lp)); std::memset(&rp, 0, sizeof(rp));
        const int L = (int)mpORBextractorLeft->mvImagePyramid.size();
        lp.nlevels = rp.nlevels = L;
        for (int l = 0; l < L && l < ORBB200_MAX_LEVELS; l++) {
            const cv::Mat& a = mpORBextractorLeft->mvImagePyramid[l];
            const cv::Mat& b = mpORBextractorRight->mvImagePyramid[l];
            lp.level[l] = a.data; lp.pitch[l] = (int32_t)a.step; lp.width[l] = a.cols; lp.height[l] = a.rows;
            rp.level[l] = b.data; rp.pitch[l] = (int32_t)b.step; rp.width[l] = b.cols; rp.height[l] = b.rows;
        }
    }
    const int L = (int)mvScaleFactors.size();
    if (orbb200_compute_stereo_matches(h, 1, &left, &right, &lp, &rp, &mvScaleFactors[0], &mvInvScaleFactors[0], L, mb, mbf,
                                       &mvuRight[0], &mvDepth[0], &nmatches, where) != ORBB200_OK) {
        std::fprintf(stderr, "Frame(B200)::ComputeStereoMatches: %s\n", orbb200_last_error());
        mvuRight.assign(N, -1.0f);
        mvDepth.assign(N, -1.0f);
    }
}

}  // namespace ORB_SLAM2
