// ref_matcher_harness.cc -- TEST INFRASTRUCTURE.  C entry points that run the reference's OWN, unmodified
// ORBmatcher.cc (with Frame.cc and MapPoint.cc) on flattened inputs: the harness builds ORB_SLAM2::Frame and
// ORB_SLAM2::MapPoint objects, calls ORBmatcher::SearchForInitialization / SearchByProjection /
// DescriptorDistance, and copies the results back.  Private members are reached by re-declaring the access
// specifiers for the reference headers only (the standard headers are included first, untouched).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <list>
#include <map>
#include <mutex>
#include <new>
#include <set>
#include <string>
#include <thread>
#include <vector>
#include <iostream>
#include <fstream>
#include <sstream>
#include <algorithm>
#include <numeric>
#include <limits>
#include <unordered_map>
#include <opencv2/opencv.hpp>
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#include <Eigen/Core>
#include <Eigen/Dense>

#define private public
#define protected public
#include "ORBmatcher.h"
#include "ORBVocabulary.h"
#undef private
#undef protected

using namespace ORB_SLAM2;

namespace {
void fill_frame(Frame& F, int n, const float* x, const float* y, const int32_t* oct, const float* ang,
                const uint8_t* desc, const float* bounds)
{
    F.N = n;
    F.mvKeysUn.resize(n);
    for (int i = 0; i < n; i++) {
        cv::KeyPoint& k = F.mvKeysUn[i];
        k.pt.x = x[i]; k.pt.y = y[i]; k.octave = oct[i]; k.angle = ang ? ang[i] : -1.f; k.size = 31.f; k.response = 0; k.class_id = -1;
    }
    F.mvKeys = F.mvKeysUn;
    F.mDescriptors = cv::Mat(n > 0 ? n : 1, 32, CV_8U, (void*)desc);
    F.mvuRight = std::vector<float>(n, -1.f);
    F.mvDepth = std::vector<float>(n, -1.f);
    F.mvpMapPoints = std::vector<MapPoint*>(n, static_cast<MapPoint*>(NULL));
    Frame::mnMinX = bounds[0]; Frame::mnMinY = bounds[1]; Frame::mnMaxX = bounds[2]; Frame::mnMaxY = bounds[3];
    // as in the Frame constructor (Frame.cc:317-318)
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
    F.AssignFeaturesToGrid();
}
}  // namespace

extern "C" {

int refm_descriptor_distance(const uint8_t* a, const uint8_t* b)
{
    cv::Mat A(1, 32, CV_8U, (void*)a), B(1, 32, CV_8U, (void*)b);
    return ORBmatcher::DescriptorDistance(A, B);
}

int refm_search_for_initialization(
    int n1, const float* k1x, const float* k1y, const int32_t* k1oct, const float* k1ang, const uint8_t* d1,
    int n2, const float* k2x, const float* k2y, const int32_t* k2oct, const float* k2ang, const uint8_t* d2,
    const float* bounds, float nnratio, int check_orientation, int window_size, float* prev_matched, int32_t* matches12)
{
    Frame F1, F2;
    fill_frame(F1, n1, k1x, k1y, k1oct, k1ang, d1, bounds);
    fill_frame(F2, n2, k2x, k2y, k2oct, k2ang, d2, bounds);
    std::vector<cv::Point2f> prev(n1);
    for (int i = 0; i < n1; i++) prev[i] = cv::Point2f(prev_matched[2 * i], prev_matched[2 * i + 1]);
    std::vector<int> m12;
    ORBmatcher matcher(nnratio, check_orientation != 0);
    const int n = matcher.SearchForInitialization(F1, F2, prev, m12, window_size);
    for (int i = 0; i < n1; i++) { matches12[i] = m12[i]; prev_matched[2 * i] = prev[i].x; prev_matched[2 * i + 1] = prev[i].y; }
    return n;
}

int refm_search_by_projection(
    int nmp, const uint8_t* mp_in_view, const uint8_t* mp_bad, const float* mp_x, const float* mp_y,
    const float* mp_xr, const int32_t* mp_level, const float* mp_viewcos, const uint8_t* mp_desc, const int32_t* mp_obs,
    int n, const float* kx, const float* ky, const int32_t* koct, const float* kuright, const uint8_t* kdesc,
    int32_t* kp_mp, const int32_t* kp_mp_obs,
    int nlevels, const float* scale_factors, const float* bounds, float nnratio, float th)
{
    Frame F;
    fill_frame(F, n, kx, ky, koct, NULL, kdesc, bounds);
    for (int i = 0; i < n; i++) F.mvuRight[i] = kuright[i];
    F.mvScaleFactors.assign(scale_factors, scale_factors + nlevels);
    // MapPoint has no default constructor and its real ones need a Map and a KeyFrame: use zeroed storage
    // and construct only what the three methods touch (descriptor Mat; the mutexes are valid when zeroed).
    MapPoint* mps = (MapPoint*)std::calloc(nmp > 0 ? nmp : 1, sizeof(MapPoint));
    MapPoint* foreign = (MapPoint*)std::calloc(n > 0 ? n : 1, sizeof(MapPoint));
    std::vector<MapPoint*> vp(nmp);
    for (int i = 0; i < nmp; i++) {
        MapPoint* p = &mps[i];
        new (&p->mDescriptor) cv::Mat(1, 32, CV_8U, (void*)(mp_desc + 32 * (size_t)i));
        p->mbTrackInView = mp_in_view[i] != 0; p->mbBad = mp_bad[i] != 0;
        p->mTrackProjX = mp_x[i]; p->mTrackProjY = mp_y[i]; p->mTrackProjXR = mp_xr[i];
        p->mnTrackScaleLevel = mp_level[i]; p->mTrackViewCos = mp_viewcos[i]; p->nObs = mp_obs[i];
        vp[i] = p;
    }
    for (int i = 0; i < n; i++) {
        if (kp_mp[i] >= 0) F.mvpMapPoints[i] = &mps[kp_mp[i]];
        else if (kp_mp[i] == -2) { foreign[i].nObs = kp_mp_obs[i]; F.mvpMapPoints[i] = &foreign[i]; }
    }
    ORBmatcher matcher(nnratio, true);
    const int cnt = matcher.SearchByProjection(F, vp, th);
    for (int i = 0; i < n; i++) {
        MapPoint* p = F.mvpMapPoints[i];
        if (!p) kp_mp[i] = -1;
        else if (p >= mps && p < mps + nmp) kp_mp[i] = (int32_t)(p - mps);
        else kp_mp[i] = -2;
    }
    for (int i = 0; i < nmp; i++) mps[i].mDescriptor.~Mat();
    std::free(mps); std::free(foreign);
    return cnt;
}

// ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (S/ORBmatcher.cc:1332-1474)
int refm_search_by_projection_last_frame(
    int nlast, const uint8_t* has_mp, const uint8_t* outlier, const float* wpos, const uint8_t* mp_desc,
    const int32_t* mp_obs, const int32_t* last_octave, const float* last_angle,
    const float* Rcw, const float* tcw, const float* K, float mbf,
    int n, const float* kx, const float* ky, const int32_t* koct, const float* kang, const float* kuright,
    const uint8_t* kdesc, int32_t* kp_mp, const int32_t* kp_mp_obs,
    int nlevels, const float* scale_factors, const float* bounds, float th, int mono, int check_orientation)
{
    Frame Cur, Last;
    fill_frame(Cur, n, kx, ky, koct, kang, kdesc, bounds);
    for (int i = 0; i < n; i++) Cur.mvuRight[i] = kuright[i];
    Cur.mvScaleFactors.assign(scale_factors, scale_factors + nlevels);
    Frame::fx = K[0]; Frame::fy = K[1]; Frame::cx = K[2]; Frame::cy = K[3];
    Cur.mbf = mbf; Cur.mb = mbf / K[0];
    float T[16] = {Rcw[0], Rcw[1], Rcw[2], tcw[0], Rcw[3], Rcw[4], Rcw[5], tcw[1], Rcw[6], Rcw[7], Rcw[8], tcw[2], 0, 0, 0, 1};
    float Tl[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    Cur.mTcw = cv::Mat(4, 4, CV_32F, T);
    Last.mTcw = cv::Mat(4, 4, CV_32F, Tl);
    // last frame: keypoints (octave from mvKeys, angle from mvKeysUn), map points, outlier flags
    std::vector<float> zx(nlast > 0 ? nlast : 1, 100.f);
    std::vector<uint8_t> zd((size_t)(nlast > 0 ? nlast : 1) * 32, 0);
    fill_frame(Last, nlast, &zx[0], &zx[0], last_octave, last_angle, &zd[0], bounds);
    fill_frame(Cur, n, kx, ky, koct, kang, kdesc, bounds);          // (the grid statics are shared: restore Cur's)
    for (int i = 0; i < n; i++) Cur.mvuRight[i] = kuright[i];
    Last.mvbOutlier = std::vector<bool>(nlast, false);
    MapPoint* mps = (MapPoint*)std::calloc(nlast > 0 ? nlast : 1, sizeof(MapPoint));
    MapPoint* foreign = (MapPoint*)std::calloc(n > 0 ? n : 1, sizeof(MapPoint));
    for (int i = 0; i < nlast; i++) {
        MapPoint* p = &mps[i];
        new (&p->mDescriptor) cv::Mat(1, 32, CV_8U, (void*)(mp_desc + 32 * (size_t)i));
        new (&p->mWorldPos) cv::Mat(3, 1, CV_32F, (void*)(wpos + 3 * (size_t)i));
        p->nObs = mp_obs[i];
        Last.mvpMapPoints[i] = has_mp[i] ? p : NULL;
        Last.mvbOutlier[i] = outlier[i] != 0;
    }
    for (int i = 0; i < n; i++) {
        if (kp_mp[i] >= 0) Cur.mvpMapPoints[i] = &mps[kp_mp[i]];
        else if (kp_mp[i] == -2) { foreign[i].nObs = kp_mp_obs[i]; Cur.mvpMapPoints[i] = &foreign[i]; }
    }
    ORBmatcher matcher(0.9f, check_orientation != 0);
    const int cnt = matcher.SearchByProjection(Cur, Last, th, mono != 0);
    for (int i = 0; i < n; i++) {
        MapPoint* p = Cur.mvpMapPoints[i];
        if (!p) kp_mp[i] = -1;
        else if (p >= mps && p < mps + nlast) kp_mp[i] = (int32_t)(p - mps);
        else kp_mp[i] = -2;
    }
    for (int i = 0; i < nlast; i++) { mps[i].mDescriptor.~Mat(); mps[i].mWorldPos.~Mat(); }
    std::free(mps); std::free(foreign);
    return cnt;
}

// ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (S/ORBmatcher.cc:1476-1603).
// valid[i] = 1: a good map point; 2: a map point that is in sAlreadyFound; 3: a bad map point; 0: empty slot.
int refm_search_by_projection_keyframe(
    int nkf, const uint8_t* valid, const float* wpos, const uint8_t* mp_desc, const float* mf_max_distance,
    const float* mf_min_distance, const float* kf_angle,
    const float* Rcw, const float* tcw, const float* K,
    int n, const float* kx, const float* ky, const int32_t* koct, const float* kang, const uint8_t* kdesc,
    int32_t* kp_mp, int nlevels, const float* scale_factors, float log_scale_factor, const float* bounds,
    float th, int orb_dist, int check_orientation, float* Ow_out)
{
    Frame Cur;
    fill_frame(Cur, n, kx, ky, koct, kang, kdesc, bounds);
    Cur.mvScaleFactors.assign(scale_factors, scale_factors + nlevels);
    Cur.mfLogScaleFactor = log_scale_factor;
    Frame::fx = K[0]; Frame::fy = K[1]; Frame::cx = K[2]; Frame::cy = K[3];
    float T[16] = {Rcw[0], Rcw[1], Rcw[2], tcw[0], Rcw[3], Rcw[4], Rcw[5], tcw[1], Rcw[6], Rcw[7], Rcw[8], tcw[2], 0, 0, 0, 1};
    Cur.mTcw = cv::Mat(4, 4, CV_32F, T);
    if (Ow_out) {       // the caller's cv::Mat expression, evaluated the way the reference evaluates it here
        const cv::Mat R = Cur.mTcw.rowRange(0, 3).colRange(0, 3), t = Cur.mTcw.rowRange(0, 3).col(3);
        const cv::Mat Ow = -R.t() * t;
        for (int r = 0; r < 3; r++) Ow_out[r] = Ow.at<float>(r);
    }
    KeyFrame* kf = (KeyFrame*)std::calloc(1, sizeof(KeyFrame));
    new (&kf->mvpMapPoints) std::vector<MapPoint*>(nkf, static_cast<MapPoint*>(NULL));
    std::vector<cv::KeyPoint>* kfKeys = const_cast<std::vector<cv::KeyPoint>*>(&kf->mvKeysUn);
    new (kfKeys) std::vector<cv::KeyPoint>(nkf);
    MapPoint* mps = (MapPoint*)std::calloc(nkf > 0 ? nkf : 1, sizeof(MapPoint));
    MapPoint* foreign = (MapPoint*)std::calloc(n > 0 ? n : 1, sizeof(MapPoint));
    std::set<MapPoint*> found;
    for (int i = 0; i < nkf; i++) {
        MapPoint* p = &mps[i];
        new (&p->mDescriptor) cv::Mat(1, 32, CV_8U, (void*)(mp_desc + 32 * (size_t)i));
        new (&p->mWorldPos) cv::Mat(3, 1, CV_32F, (void*)(wpos + 3 * (size_t)i));
        p->mfMaxDistance = mf_max_distance[i]; p->mfMinDistance = mf_min_distance[i];
        p->mbBad = valid[i] == 3;
        if (valid[i] == 2) found.insert(p);
        kf->mvpMapPoints[i] = valid[i] ? p : NULL;
        (*kfKeys)[i].angle = kf_angle[i];
    }
    for (int i = 0; i < n; i++) {
        if (kp_mp[i] >= 0) Cur.mvpMapPoints[i] = &mps[kp_mp[i]];
        else if (kp_mp[i] != -1) Cur.mvpMapPoints[i] = &foreign[i];
    }
    ORBmatcher matcher(0.9f, check_orientation != 0);
    const int cnt = matcher.SearchByProjection(Cur, kf, found, th, orb_dist);
    for (int i = 0; i < n; i++) {
        MapPoint* p = Cur.mvpMapPoints[i];
        if (!p) kp_mp[i] = -1;
        else if (p >= mps && p < mps + nkf) kp_mp[i] = (int32_t)(p - mps);
        else kp_mp[i] = -2;
    }
    for (int i = 0; i < nkf; i++) { mps[i].mDescriptor.~Mat(); mps[i].mWorldPos.~Mat(); }
    kf->mvpMapPoints.~vector(); kfKeys->~vector();
    std::free(mps); std::free(foreign); std::free(kf);
    return cnt;
}

// ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) (S/ORBmatcher.cc:161-292).
// kf_valid[i]: 0 = empty slot, 1 = good map point, 3 = bad map point.  Feature vectors arrive flattened and are
// rebuilt with DBoW2::FeatureVector::addFeature (the reference's own Thirdparty/DBoW2/src/FeatureVector.cpp).
int refm_search_by_bow(
    int nkf, const uint8_t* kf_valid, const uint8_t* kf_desc, const float* kf_angle,
    int kf_nn, const uint32_t* kf_node, const int32_t* kf_start, const uint32_t* kf_feat,
    int nf, const uint8_t* f_desc, const float* f_angle,
    int f_nn, const uint32_t* f_node, const int32_t* f_start, const uint32_t* f_feat,
    float nnratio, int check_orientation, int32_t* matches)
{
    Frame F;
    F.N = nf;
    F.mvKeys.resize(nf);
    for (int i = 0; i < nf; i++) F.mvKeys[i].angle = f_angle[i];
    F.mDescriptors = cv::Mat(nf > 0 ? nf : 1, 32, CV_8U, (void*)f_desc);
    for (int a = 0; a < f_nn; a++)
        for (int j = f_start[a]; j < f_start[a + 1]; j++) F.mFeatVec.addFeature(f_node[a], f_feat[j]);

    KeyFrame* kf = (KeyFrame*)std::calloc(1, sizeof(KeyFrame));
    new (&kf->mvpMapPoints) std::vector<MapPoint*>(nkf, static_cast<MapPoint*>(NULL));
    std::vector<cv::KeyPoint>* kfKeys = const_cast<std::vector<cv::KeyPoint>*>(&kf->mvKeysUn);
    new (kfKeys) std::vector<cv::KeyPoint>(nkf);
    cv::Mat* kfDesc = const_cast<cv::Mat*>(&kf->mDescriptors);
    new (kfDesc) cv::Mat(nkf > 0 ? nkf : 1, 32, CV_8U, (void*)kf_desc);
    new (&kf->mFeatVec) DBoW2::FeatureVector();
    for (int a = 0; a < kf_nn; a++)
        for (int j = kf_start[a]; j < kf_start[a + 1]; j++) kf->mFeatVec.addFeature(kf_node[a], kf_feat[j]);
    MapPoint* mps = (MapPoint*)std::calloc(nkf > 0 ? nkf : 1, sizeof(MapPoint));
    for (int i = 0; i < nkf; i++) {
        mps[i].mbBad = kf_valid[i] == 3;
        kf->mvpMapPoints[i] = kf_valid[i] ? &mps[i] : NULL;
        (*kfKeys)[i].angle = kf_angle[i];
    }
    ORBmatcher matcher(nnratio, check_orientation != 0);
    std::vector<MapPoint*> out;
    const int cnt = matcher.SearchByBoW(kf, F, out);
    for (int i = 0; i < nf; i++) matches[i] = out[i] ? (int32_t)(out[i] - mps) : -1;
    kf->mFeatVec.~FeatureVector(); kfDesc->~Mat(); kf->mvpMapPoints.~vector(); kfKeys->~vector();
    std::free(mps); std::free(kf);
    return cnt;
}

namespace {
struct FakeKeyFrame {
    KeyFrame* kf; MapPoint* mps; int n;
    FakeKeyFrame(int n_, const uint8_t* valid, const uint8_t* desc, const float* angle, int nn, const uint32_t* node,
                 const int32_t* start, const uint32_t* feat) : n(n_)
    {
        kf = (KeyFrame*)std::calloc(1, sizeof(KeyFrame));
        new (&kf->mvpMapPoints) std::vector<MapPoint*>(n, static_cast<MapPoint*>(NULL));
        new (const_cast<std::vector<cv::KeyPoint>*>(&kf->mvKeysUn)) std::vector<cv::KeyPoint>(n);
        new (const_cast<cv::Mat*>(&kf->mDescriptors)) cv::Mat(n > 0 ? n : 1, 32, CV_8U, (void*)desc);
        new (&kf->mFeatVec) DBoW2::FeatureVector();
        for (int a = 0; a < nn; a++)
            for (int j = start[a]; j < start[a + 1]; j++) kf->mFeatVec.addFeature(node[a], feat[j]);
        mps = (MapPoint*)std::calloc(n > 0 ? n : 1, sizeof(MapPoint));
        for (int i = 0; i < n; i++) {
            mps[i].mbBad = valid[i] == 3;
            kf->mvpMapPoints[i] = valid[i] ? &mps[i] : NULL;
            const_cast<std::vector<cv::KeyPoint>&>(kf->mvKeysUn)[i].angle = angle[i];
        }
    }
    ~FakeKeyFrame()
    {
        kf->mFeatVec.~FeatureVector(); const_cast<cv::Mat*>(&kf->mDescriptors)->~Mat(); kf->mvpMapPoints.~vector();
        const_cast<std::vector<cv::KeyPoint>*>(&kf->mvKeysUn)->~vector();
        std::free(mps); std::free(kf);
    }
};
}  // namespace

// ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12) (S/ORBmatcher.cc:526-659)
int refm_search_by_bow_keyframes(
    int n1, const uint8_t* valid1, const uint8_t* desc1, const float* angle1,
    int nn1, const uint32_t* node1, const int32_t* start1, const uint32_t* feat1,
    int n2, const uint8_t* valid2, const uint8_t* desc2, const float* angle2,
    int nn2, const uint32_t* node2, const int32_t* start2, const uint32_t* feat2,
    float nnratio, int check_orientation, int32_t* matches12)
{
    FakeKeyFrame k1(n1, valid1, desc1, angle1, nn1, node1, start1, feat1), k2(n2, valid2, desc2, angle2, nn2, node2, start2, feat2);
    ORBmatcher matcher(nnratio, check_orientation != 0);
    std::vector<MapPoint*> out;
    const int cnt = matcher.SearchByBoW(k1.kf, k2.kf, out);
    for (int i = 0; i < n1; i++) matches12[i] = out[i] ? (int32_t)(out[i] - k2.mps) : -1;
    return cnt;
}

// ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) (S/ORBmatcher.cc:661-827).
// has_mp: the slot holds a map point.  pose2 = {R2w (9, row major), t2w (3)}, Cw1 = pKF1->GetCameraCenter(),
// K2 = {fx, fy, cx, cy}.  epipole_out = {ex, ey} as the same expressions give them here (:668-675).
int refm_search_for_triangulation(
    int n1, const uint8_t* has_mp1, const uint8_t* desc1, const float* x1, const float* y1, const float* angle1, const float* uright1,
    int nn1, const uint32_t* node1, const int32_t* start1, const uint32_t* feat1,
    int n2, const uint8_t* has_mp2, const uint8_t* desc2, const float* x2, const float* y2, const int32_t* oct2, const float* angle2,
    const float* uright2, int nn2, const uint32_t* node2, const int32_t* start2, const uint32_t* feat2,
    const float* F12, const float* Cw1, const float* pose2, const float* K2, int nlevels, const float* scale_factors2,
    const float* level_sigma2_2, int only_stereo, int check_orientation, int32_t* matches12, float* epipole_out)
{
    FakeKeyFrame k1(n1, has_mp1, desc1, angle1, nn1, node1, start1, feat1), k2(n2, has_mp2, desc2, angle2, nn2, node2, start2, feat2);
    struct Fill {
        static void run(KeyFrame* kf, int n, const float* x, const float* y, const int32_t* oct, const float* ur)
        {
            *const_cast<int*>(&kf->N) = n;
            std::vector<cv::KeyPoint>& k = const_cast<std::vector<cv::KeyPoint>&>(kf->mvKeysUn);
            for (int i = 0; i < n; i++) { k[i].pt.x = x[i]; k[i].pt.y = y[i]; k[i].octave = oct ? oct[i] : 0; }
            new (const_cast<std::vector<float>*>(&kf->mvuRight)) std::vector<float>(ur, ur + n);
        }
    };
    Fill::run(k1.kf, n1, x1, y1, NULL, uright1);
    Fill::run(k2.kf, n2, x2, y2, oct2, uright2);
    new (const_cast<std::vector<float>*>(&k2.kf->mvScaleFactors)) std::vector<float>(scale_factors2, scale_factors2 + nlevels);
    new (const_cast<std::vector<float>*>(&k2.kf->mvLevelSigma2)) std::vector<float>(level_sigma2_2, level_sigma2_2 + nlevels);
    *const_cast<float*>(&k2.kf->fx) = K2[0]; *const_cast<float*>(&k2.kf->fy) = K2[1];
    *const_cast<float*>(&k2.kf->cx) = K2[2]; *const_cast<float*>(&k2.kf->cy) = K2[3];
    float T2[16] = {pose2[0], pose2[1], pose2[2], pose2[9], pose2[3], pose2[4], pose2[5], pose2[10], pose2[6], pose2[7], pose2[8], pose2[11], 0, 0, 0, 1};
    float C1[3] = {Cw1[0], Cw1[1], Cw1[2]};
    new (&k2.kf->Tcw) cv::Mat(4, 4, CV_32F, T2);
    new (&k1.kf->Ow) cv::Mat(3, 1, CV_32F, C1);
    cv::Mat F(3, 3, CV_32F, (void*)F12);
    if (epipole_out) {
        cv::Mat Cw = k1.kf->GetCameraCenter();
        cv::Mat R2w = k2.kf->GetRotation();
        cv::Mat t2w = k2.kf->GetTranslation();
        cv::Mat C2 = R2w * Cw + t2w;
        const float invz = 1.0f / C2.at<float>(2);
        epipole_out[0] = k2.kf->fx * C2.at<float>(0) * invz + k2.kf->cx;
        epipole_out[1] = k2.kf->fy * C2.at<float>(1) * invz + k2.kf->cy;
    }
    ORBmatcher matcher(0.6f, check_orientation != 0);
    std::vector<std::pair<size_t, size_t> > pairs;
    const int cnt = matcher.SearchForTriangulation(k1.kf, k2.kf, F, pairs, only_stereo != 0);
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    for (size_t k = 0; k < pairs.size(); k++) matches12[pairs[k].first] = (int32_t)pairs[k].second;
    k2.kf->Tcw.~Mat(); k1.kf->Ow.~Mat();
    const_cast<std::vector<float>*>(&k2.kf->mvScaleFactors)->~vector(); const_cast<std::vector<float>*>(&k2.kf->mvLevelSigma2)->~vector();
    const_cast<std::vector<float>*>(&k1.kf->mvuRight)->~vector(); const_cast<std::vector<float>*>(&k2.kf->mvuRight)->~vector();
    return cnt;
}

// MapPoint::ComputeDistinctiveDescriptors (S/MapPoint.cc:248-313): nobs observing key frames (allocated as one array,
// so the std::map<KeyFrame*, size_t> iterates them in index order), key frame k observes row idx[k] of its own
// descriptor matrix (a 4-row matrix whose row idx[k] is desc[k]); bad[k] marks key frames that are skipped.
// Writes the chosen descriptor to out32; returns 1 if the map point's descriptor was set.
int refm_compute_distinctive_descriptors(int nobs, const uint8_t* desc, const uint8_t* bad, uint8_t* out32)
{
    MapPoint* mp = (MapPoint*)std::calloc(1, sizeof(MapPoint));
    new (&mp->mObservations) std::map<KeyFrame*, size_t>();
    new (&mp->mDescriptor) cv::Mat();
    KeyFrame* kfs = (KeyFrame*)std::calloc(nobs > 0 ? nobs : 1, sizeof(KeyFrame));
    std::vector<std::vector<uint8_t> > store(nobs, std::vector<uint8_t>(4 * 32, 0));
    for (int k = 0; k < nobs; k++) {
        const size_t row = (size_t)(k % 4);
        std::memcpy(&store[k][row * 32], desc + 32 * (size_t)k, 32);
        new (const_cast<cv::Mat*>(&kfs[k].mDescriptors)) cv::Mat(4, 32, CV_8U, &store[k][0]);
        kfs[k].mbBad = bad[k] != 0;
        mp->mObservations[&kfs[k]] = row;
    }
    mp->ComputeDistinctiveDescriptors();
    const int set = !mp->mDescriptor.empty();
    if (set) std::memcpy(out32, mp->mDescriptor.ptr<uint8_t>(), 32);
    for (int k = 0; k < nobs; k++) const_cast<cv::Mat*>(&kfs[k].mDescriptors)->~Mat();
    mp->mDescriptor.~Mat(); mp->mObservations.~map();
    std::free(kfs); std::free(mp);
    return set;
}

// ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, th) (S/ORBmatcher.cc:829-975), called once per
// candidate map point on a key frame whose keypoints hold no map points, so the call ends in the "add observation"
// branch and the chosen keypoint can be read back from pMP->mObservations: this pins the search, which is the part
// that runs on the device (the replace-or-add surgery stays host code in the shim).
// valid: 1 = usable, 2 = already observed by the key frame (IsInKeyFrame), 3 = bad, 0 = NULL.
void refm_fuse_search(
    int nmp, const uint8_t* valid, const float* wpos, const float* normal, const uint8_t* mp_desc,
    const float* mf_max_distance, const float* mf_min_distance,
    const float* Rcw, const float* tcw, const float* Ow, const float* K, float bf,
    int n, const float* kx, const float* ky, const int32_t* koct, const float* kuright, const uint8_t* kdesc,
    int nlevels, const float* scale_factors, const float* inv_level_sigma2, float log_scale_factor,
    const float* bounds, float th, int32_t* best_idx)
{
    // the Frame the key frame was made from: its grid (float bounds) is what KeyFrame::mGrid holds
    Frame F;
    fill_frame(F, n, kx, ky, koct, NULL, kdesc, bounds);
    KeyFrame* kf = (KeyFrame*)std::calloc(1, sizeof(KeyFrame));
    *const_cast<int*>(&kf->N) = n;
    *const_cast<int*>(&kf->mnGridCols) = FRAME_GRID_COLS; *const_cast<int*>(&kf->mnGridRows) = FRAME_GRID_ROWS;
    *const_cast<float*>(&kf->mfGridElementWidthInv) = Frame::mfGridElementWidthInv;
    *const_cast<float*>(&kf->mfGridElementHeightInv) = Frame::mfGridElementHeightInv;
    *const_cast<int*>(&kf->mnMinX) = Frame::mnMinX; *const_cast<int*>(&kf->mnMinY) = Frame::mnMinY;        // float -> int as in KeyFrame.cc:42
    *const_cast<int*>(&kf->mnMaxX) = Frame::mnMaxX; *const_cast<int*>(&kf->mnMaxY) = Frame::mnMaxY;
    *const_cast<float*>(&kf->fx) = K[0]; *const_cast<float*>(&kf->fy) = K[1]; *const_cast<float*>(&kf->cx) = K[2]; *const_cast<float*>(&kf->cy) = K[3];
    *const_cast<float*>(&kf->mbf) = bf; *const_cast<float*>(&kf->mfLogScaleFactor) = log_scale_factor;
    new (&kf->mGrid) std::vector<std::vector<std::vector<size_t> > >(FRAME_GRID_COLS);
    for (int i = 0; i < FRAME_GRID_COLS; i++) { kf->mGrid[i].resize(FRAME_GRID_ROWS); for (int j = 0; j < FRAME_GRID_ROWS; j++) kf->mGrid[i][j] = F.mGrid[i][j]; }
    new (const_cast<std::vector<cv::KeyPoint>*>(&kf->mvKeysUn)) std::vector<cv::KeyPoint>(F.mvKeysUn);
    new (const_cast<std::vector<float>*>(&kf->mvuRight)) std::vector<float>(kuright, kuright + n);
    new (const_cast<cv::Mat*>(&kf->mDescriptors)) cv::Mat(n > 0 ? n : 1, 32, CV_8U, (void*)kdesc);
    new (const_cast<std::vector<float>*>(&kf->mvScaleFactors)) std::vector<float>(scale_factors, scale_factors + nlevels);
    new (const_cast<std::vector<float>*>(&kf->mvInvLevelSigma2)) std::vector<float>(inv_level_sigma2, inv_level_sigma2 + nlevels);
    new (&kf->mvpMapPoints) std::vector<MapPoint*>(n, static_cast<MapPoint*>(NULL));
    float T[16] = {Rcw[0], Rcw[1], Rcw[2], tcw[0], Rcw[3], Rcw[4], Rcw[5], tcw[1], Rcw[6], Rcw[7], Rcw[8], tcw[2], 0, 0, 0, 1};
    float O3[3] = {Ow[0], Ow[1], Ow[2]};
    new (&kf->Tcw) cv::Mat(4, 4, CV_32F, T);
    new (&kf->Ow) cv::Mat(3, 1, CV_32F, O3);

    ORBmatcher matcher(0.6f, true);
    MapPoint* mps = (MapPoint*)std::calloc(nmp > 0 ? nmp : 1, sizeof(MapPoint));
    for (int i = 0; i < nmp; i++) {
        best_idx[i] = -1;
        MapPoint* p = &mps[i];
        new (&p->mDescriptor) cv::Mat(1, 32, CV_8U, (void*)(mp_desc + 32 * (size_t)i));
        new (&p->mWorldPos) cv::Mat(3, 1, CV_32F, (void*)(wpos + 3 * (size_t)i));
        new (&p->mNormalVector) cv::Mat(3, 1, CV_32F, (void*)(normal + 3 * (size_t)i));
        new (&p->mObservations) std::map<KeyFrame*, size_t>();
        p->mfMaxDistance = mf_max_distance[i]; p->mfMinDistance = mf_min_distance[i];
        p->mbBad = valid[i] == 3;
        if (valid[i] == 2) p->mObservations[kf] = 0;
        std::vector<MapPoint*> one(1, valid[i] ? p : static_cast<MapPoint*>(NULL));
        const int fused = matcher.Fuse(kf, one, th);
        if (fused && valid[i] == 1) {
            best_idx[i] = (int32_t)p->mObservations[kf];
            kf->mvpMapPoints[best_idx[i]] = NULL;                 // next map point sees an empty key frame again
        }
        p->mObservations.~map(); p->mNormalVector.~Mat(); p->mWorldPos.~Mat(); p->mDescriptor.~Mat();
    }
    std::free(mps);
    kf->Ow.~Mat(); kf->Tcw.~Mat(); kf->mvpMapPoints.~vector(); kf->mGrid.~vector();
    const_cast<std::vector<float>*>(&kf->mvInvLevelSigma2)->~vector(); const_cast<std::vector<float>*>(&kf->mvScaleFactors)->~vector();
    const_cast<cv::Mat*>(&kf->mDescriptors)->~Mat(); const_cast<std::vector<float>*>(&kf->mvuRight)->~vector();
    const_cast<std::vector<cv::KeyPoint>*>(&kf->mvKeysUn)->~vector();
    std::free(kf);
}

namespace {
// a key frame with everything the projection searches read: keypoints, descriptors, the Frame-assigned grid, the
// int-truncated bounds, intrinsics, pyramid tables and pose
struct FullKeyFrame {
    KeyFrame* kf; MapPoint* mps; int n; float T[16], O3[3];
    FullKeyFrame(int n_, const float* kx, const float* ky, const int32_t* koct, const float* kuright, const uint8_t* kdesc,
                 const float* K, float bf, int nlevels, const float* sf, const float* ils, float logsf, const float* bounds,
                 const float* Rcw, const float* tcw, const float* Ow) : n(n_)
    {
        Frame F;
        fill_frame(F, n, kx, ky, koct, NULL, kdesc, bounds);
        kf = (KeyFrame*)std::calloc(1, sizeof(KeyFrame));
        *const_cast<int*>(&kf->N) = n;
        *const_cast<int*>(&kf->mnGridCols) = FRAME_GRID_COLS; *const_cast<int*>(&kf->mnGridRows) = FRAME_GRID_ROWS;
        *const_cast<float*>(&kf->mfGridElementWidthInv) = Frame::mfGridElementWidthInv;
        *const_cast<float*>(&kf->mfGridElementHeightInv) = Frame::mfGridElementHeightInv;
        *const_cast<int*>(&kf->mnMinX) = Frame::mnMinX; *const_cast<int*>(&kf->mnMinY) = Frame::mnMinY;
        *const_cast<int*>(&kf->mnMaxX) = Frame::mnMaxX; *const_cast<int*>(&kf->mnMaxY) = Frame::mnMaxY;
        *const_cast<float*>(&kf->fx) = K[0]; *const_cast<float*>(&kf->fy) = K[1]; *const_cast<float*>(&kf->cx) = K[2]; *const_cast<float*>(&kf->cy) = K[3];
        *const_cast<float*>(&kf->mbf) = bf; *const_cast<float*>(&kf->mfLogScaleFactor) = logsf;
        new (&kf->mGrid) std::vector<std::vector<std::vector<size_t> > >(FRAME_GRID_COLS);
        for (int i = 0; i < FRAME_GRID_COLS; i++) { kf->mGrid[i].resize(FRAME_GRID_ROWS); for (int j = 0; j < FRAME_GRID_ROWS; j++) kf->mGrid[i][j] = F.mGrid[i][j]; }
        new (const_cast<std::vector<cv::KeyPoint>*>(&kf->mvKeysUn)) std::vector<cv::KeyPoint>(F.mvKeysUn);
        new (const_cast<std::vector<float>*>(&kf->mvuRight)) std::vector<float>(kuright, kuright + n);
        new (const_cast<cv::Mat*>(&kf->mDescriptors)) cv::Mat(n > 0 ? n : 1, 32, CV_8U, (void*)kdesc);
        new (const_cast<std::vector<float>*>(&kf->mvScaleFactors)) std::vector<float>(sf, sf + nlevels);
        new (const_cast<std::vector<float>*>(&kf->mvInvLevelSigma2)) std::vector<float>(ils, ils + nlevels);
        new (&kf->mvpMapPoints) std::vector<MapPoint*>(n, static_cast<MapPoint*>(NULL));
        const float t[16] = {Rcw[0], Rcw[1], Rcw[2], tcw[0], Rcw[3], Rcw[4], Rcw[5], tcw[1], Rcw[6], Rcw[7], Rcw[8], tcw[2], 0, 0, 0, 1};
        std::memcpy(T, t, sizeof(T));
        O3[0] = Ow[0]; O3[1] = Ow[1]; O3[2] = Ow[2];
        new (&kf->Tcw) cv::Mat(4, 4, CV_32F, T);
        new (&kf->Ow) cv::Mat(3, 1, CV_32F, O3);
        mps = NULL;
    }
    ~FullKeyFrame()
    {
        kf->Ow.~Mat(); kf->Tcw.~Mat(); kf->mvpMapPoints.~vector(); kf->mGrid.~vector();
        const_cast<std::vector<float>*>(&kf->mvInvLevelSigma2)->~vector(); const_cast<std::vector<float>*>(&kf->mvScaleFactors)->~vector();
        const_cast<cv::Mat*>(&kf->mDescriptors)->~Mat(); const_cast<std::vector<float>*>(&kf->mvuRight)->~vector();
        const_cast<std::vector<cv::KeyPoint>*>(&kf->mvKeysUn)->~vector();
        std::free(kf);
    }
};

struct FakePoints {
    MapPoint* mps; int n;
    FakePoints(int n_, const uint8_t* valid, const float* wpos, const float* normal, const uint8_t* desc, const float* mx, const float* mn) : n(n_)
    {
        mps = (MapPoint*)std::calloc(n > 0 ? n : 1, sizeof(MapPoint));
        for (int i = 0; i < n; i++) {
            MapPoint* p = &mps[i];
            new (&p->mDescriptor) cv::Mat(1, 32, CV_8U, (void*)(desc + 32 * (size_t)i));
            new (&p->mWorldPos) cv::Mat(3, 1, CV_32F, (void*)(wpos + 3 * (size_t)i));
            new (&p->mNormalVector) cv::Mat(3, 1, CV_32F, (void*)(normal + 3 * (size_t)i));
            new (&p->mObservations) std::map<KeyFrame*, size_t>();
            p->mfMaxDistance = mx[i]; p->mfMinDistance = mn[i];
            p->mbBad = valid[i] == 3;
            p->mnId = i + 1;
        }
    }
    ~FakePoints()
    {
        for (int i = 0; i < n; i++) { MapPoint* p = &mps[i]; p->mObservations.~map(); p->mNormalVector.~Mat(); p->mWorldPos.~Mat(); p->mDescriptor.~Mat(); }
        std::free(mps);
    }
};
}  // namespace

// ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>& vpReplacePoint)
// (S/ORBmatcher.cc:979-1104), one candidate per call on an empty key frame (see refm_fuse_search).  Scw = [Rcw | tcw]
// with unit scale, so the decomposition of :987-991 returns the inputs.  valid: 1 usable, 3 bad, anything else is not
// passed (this overload takes no NULL candidates and an empty key frame has nothing "already found"); Ow_out = the camera centre as the reference's expression evaluates it here.
void refm_fuse_search_sim3(
    int nmp, const uint8_t* valid, const float* wpos, const float* normal, const uint8_t* mp_desc,
    const float* mf_max_distance, const float* mf_min_distance,
    const float* Rcw, const float* tcw, const float* Ow, const float* K, float bf,
    int n, const float* kx, const float* ky, const int32_t* koct, const float* kuright, const uint8_t* kdesc,
    int nlevels, const float* scale_factors, const float* inv_level_sigma2, float log_scale_factor,
    const float* bounds, float th, int32_t* best_idx)
{
    FullKeyFrame k(n, kx, ky, koct, kuright, kdesc, K, bf, nlevels, scale_factors, inv_level_sigma2, log_scale_factor, bounds, Rcw, tcw, Ow);
    FakePoints pts(nmp, valid, wpos, normal, mp_desc, mf_max_distance, mf_min_distance);
    cv::Mat Scw(4, 4, CV_32F, k.T);
    ORBmatcher matcher(0.6f, true);
    for (int i = 0; i < nmp; i++) {
        best_idx[i] = -1;
        if (valid[i] != 1 && valid[i] != 3) continue;
        std::vector<MapPoint*> one(1, &pts.mps[i]), rep(1, static_cast<MapPoint*>(NULL));
        const int fused = matcher.Fuse(k.kf, Scw, one, th, rep);
        if (fused) {
            best_idx[i] = (int32_t)pts.mps[i].mObservations[k.kf];
            k.kf->mvpMapPoints[best_idx[i]] = NULL;
        }
    }
}

// ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th) (S/ORBmatcher.cc:1106-1330) with s12 = 1 and no
// previous matches.  Key frame k holds map point i at slot i (valid: 0 = empty slot, 1 good, 3 bad).
// matches12[n1] out = slot of key frame 2 or -1.  Returns nFound.
int refm_search_by_sim3(
    int n1, const uint8_t* valid1, const float* wpos1, const uint8_t* mpdesc1, const float* mx1, const float* mn1,
    const float* kx1, const float* ky1, const int32_t* koct1, const uint8_t* kdesc1, const float* Rcw1, const float* tcw1,
    int n2, const uint8_t* valid2, const float* wpos2, const uint8_t* mpdesc2, const float* mx2, const float* mn2,
    const float* kx2, const float* ky2, const int32_t* koct2, const uint8_t* kdesc2, const float* Rcw2, const float* tcw2,
    const float* K, int nlevels, const float* scale_factors, const float* inv_level_sigma2, float log_scale_factor,
    const float* bounds, const float* R12, const float* t12, float th, int32_t* matches12)
{
    std::vector<float> ur1(n1 > 0 ? n1 : 1, -1.f), ur2(n2 > 0 ? n2 : 1, -1.f), zero(3, 0.f);
    std::vector<float> nrm1((size_t)(n1 > 0 ? n1 : 1) * 3, 0.f), nrm2((size_t)(n2 > 0 ? n2 : 1) * 3, 0.f);
    FullKeyFrame k2(n2, kx2, ky2, koct2, &ur2[0], kdesc2, K, 0.f, nlevels, scale_factors, inv_level_sigma2, log_scale_factor, bounds, Rcw2, tcw2, &zero[0]);
    FullKeyFrame k1(n1, kx1, ky1, koct1, &ur1[0], kdesc1, K, 0.f, nlevels, scale_factors, inv_level_sigma2, log_scale_factor, bounds, Rcw1, tcw1, &zero[0]);
    FakePoints p1(n1, valid1, wpos1, &nrm1[0], mpdesc1, mx1, mn1), p2(n2, valid2, wpos2, &nrm2[0], mpdesc2, mx2, mn2);
    for (int i = 0; i < n1; i++) k1.kf->mvpMapPoints[i] = valid1[i] ? &p1.mps[i] : NULL;
    for (int i = 0; i < n2; i++) k2.kf->mvpMapPoints[i] = valid2[i] ? &p2.mps[i] : NULL;
    float Rm[9], tm[3];
    std::memcpy(Rm, R12, sizeof(Rm)); std::memcpy(tm, t12, sizeof(tm));
    cv::Mat R(3, 3, CV_32F, Rm), t(3, 1, CV_32F, tm);
    std::vector<MapPoint*> m12(n1, static_cast<MapPoint*>(NULL));
    ORBmatcher matcher(0.75f, true);
    const float s12 = 1.0f;
    const int found = matcher.SearchBySim3(k1.kf, k2.kf, m12, s12, R, t, th);
    for (int i = 0; i < n1; i++) matches12[i] = m12[i] ? (int32_t)(m12[i] - p2.mps) : -1;
    return found;
}

// ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched,
// int th) (S/ORBmatcher.cc:294-407), Scw = [Rcw | tcw] with unit scale.  valid: 1 usable, 3 bad; candidates with any
// other code are not in the list the reference sees.  matched[n] in/out as in the oracle.
int refm_search_by_projection_sim3(
    int nmp, const uint8_t* valid, const float* wpos, const float* normal, const uint8_t* mp_desc,
    const float* mf_max_distance, const float* mf_min_distance,
    const float* Rcw, const float* tcw, const float* Ow, const float* K,
    int n, const float* kx, const float* ky, const int32_t* koct, const uint8_t* kdesc,
    int nlevels, const float* scale_factors, float log_scale_factor, const float* bounds, int th, int32_t* matched)
{
    std::vector<float> ur(n > 0 ? n : 1, -1.f), ils(nlevels, 1.f);
    FullKeyFrame k(n, kx, ky, koct, &ur[0], kdesc, K, 0.f, nlevels, scale_factors, &ils[0], log_scale_factor, bounds, Rcw, tcw, Ow);
    FakePoints pts(nmp, valid, wpos, normal, mp_desc, mf_max_distance, mf_min_distance);
    MapPoint* foreign = (MapPoint*)std::calloc(n > 0 ? n : 1, sizeof(MapPoint));
    std::vector<MapPoint*> list, vpMatched(n, static_cast<MapPoint*>(NULL));
    std::vector<int> listIdx;
    for (int i = 0; i < nmp; i++) if (valid[i] == 1 || valid[i] == 3) { list.push_back(&pts.mps[i]); listIdx.push_back(i); }
    for (int i = 0; i < n; i++) {
        if (matched[i] >= 0) vpMatched[i] = &pts.mps[matched[i]];
        else if (matched[i] != -1) vpMatched[i] = &foreign[i];
    }
    cv::Mat Scw(4, 4, CV_32F, k.T);
    ORBmatcher matcher(0.75f, true);
    const int cnt = matcher.SearchByProjection(k.kf, Scw, list, vpMatched, th);
    for (int i = 0; i < n; i++) {
        MapPoint* p = vpMatched[i];
        if (!p) matched[i] = -1;
        else if (p >= pts.mps && p < pts.mps + nmp) matched[i] = (int32_t)(p - pts.mps);
        else matched[i] = -2;
    }
    std::free(foreign);
    return cnt;
}

// DBoW2 ORBVocabulary::transform(features, BowVector, FeatureVector, levelsup) (Frame::ComputeBoW, S/Frame.cc:520-527)
// on a vocabulary loaded by the reference's own loadFromTextFile from `path`.  Outputs flattened in map order.
// Returns 0, or -1 if the file did not load.
int refm_bow_transform(const char* path, int n, const uint8_t* desc, int levelsup,
                       int32_t* bow_n, uint32_t* bow_word, double* bow_value,
                       int32_t* fv_n, uint32_t* fv_node, int32_t* fv_start, uint32_t* fv_feat)
{
    static std::map<std::string, ORBVocabulary*> cache;
    ORBVocabulary*& voc = cache[path];
    if (!voc) {
        voc = new ORBVocabulary();
        if (!voc->loadFromTextFile(path)) { delete voc; voc = NULL; cache.erase(path); return -1; }
    }
    cv::Mat D(n > 0 ? n : 1, 32, CV_8U, (void*)desc);
    std::vector<cv::Mat> feats;
    for (int j = 0; j < n; j++) feats.push_back(D.row(j));                 // Converter::toDescriptorVector
    DBoW2::BowVector bv;
    DBoW2::FeatureVector fv;
    voc->transform(feats, bv, fv, levelsup);
    int k = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++k) { bow_word[k] = it->first; bow_value[k] = it->second; }
    *bow_n = k;
    int a = 0, pos = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++a) {
        fv_node[a] = it->first; fv_start[a] = pos;
        for (size_t j = 0; j < it->second.size(); j++) fv_feat[pos++] = it->second[j];
    }
    fv_start[a] = pos;
    *fv_n = a;
    return 0;
}

// Frame::ComputeStereoMatches (S/Frame.cc:591-763) on a Frame whose two extractors only carry their image pyramids.
// An ORBextractor cannot be constructed here (ORBextractor.cc is not part of this library): zeroed storage with the
// one member the function reads, mvImagePyramid, constructed in place.
int refm_compute_stereo_matches(
    int n, const float* kx, const float* ky, const int32_t* koct, const uint8_t* desc,
    int nr, const float* rx, const float* ry, const int32_t* roct, const uint8_t* rdesc,
    int nlevels, const float* scale, const float* inv_scale,
    const uint8_t* const* limg, const int32_t* lpitch, const uint8_t* const* rimg, const int32_t* rpitch,
    const int32_t* lw, const int32_t* lh, float mb, float mbf, float* u_right, float* depth)
{
    Frame F;
    F.N = n;
    F.mvKeys.resize(n); F.mvKeysRight.resize(nr);
    for (int i = 0; i < n; i++) { cv::KeyPoint& k = F.mvKeys[i]; k.pt.x = kx[i]; k.pt.y = ky[i]; k.octave = koct[i]; k.angle = 0; k.size = 31.f; }
    for (int i = 0; i < nr; i++) { cv::KeyPoint& k = F.mvKeysRight[i]; k.pt.x = rx[i]; k.pt.y = ry[i]; k.octave = roct[i]; k.angle = 0; k.size = 31.f; }
    F.mDescriptors = cv::Mat(n > 0 ? n : 1, 32, CV_8U, (void*)desc);
    F.mDescriptorsRight = cv::Mat(nr > 0 ? nr : 1, 32, CV_8U, (void*)rdesc);
    F.mvScaleFactors.assign(scale, scale + nlevels);
    F.mvInvScaleFactors.assign(inv_scale, inv_scale + nlevels);
    F.mb = mb; F.mbf = mbf;
    ORBextractor* ex[2];
    for (int s = 0; s < 2; s++) {
        ex[s] = (ORBextractor*)std::calloc(1, sizeof(ORBextractor));
        new (&ex[s]->mvImagePyramid) std::vector<cv::Mat>(nlevels);
        for (int l = 0; l < nlevels; l++)
            ex[s]->mvImagePyramid[l] = cv::Mat(lh[l], lw[l], CV_8U, (void*)(s ? rimg[l] : limg[l]), (size_t)(s ? rpitch[l] : lpitch[l]));
    }
    F.mpORBextractorLeft = ex[0]; F.mpORBextractorRight = ex[1];
    F.ComputeStereoMatches();
    int cnt = 0;
    for (int i = 0; i < n; i++) { u_right[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; cnt += F.mvuRight[i] != -1.0f; }
    for (int s = 0; s < 2; s++) { ex[s]->mvImagePyramid.~vector(); std::free(ex[s]); }
    F.mpORBextractorLeft = F.mpORBextractorRight = NULL;
    return cnt;
}
}
