// ORBmatcher_b200.cc -- B200 bodies for the search routines of ORB_SLAM2::ORBmatcher (the three core ones first):
//   DescriptorDistance                                   (replaces S/ORBmatcher.cc:1651-1667)
//   SearchForInitialization                              (replaces S/ORBmatcher.cc:409-524)
//   SearchByProjection(Frame&, vector<MapPoint*>&, th)   (replaces S/ORBmatcher.cc:47-131)
//   SearchByProjection(CurrentFrame, LastFrame, th, bMono) (replaces S/ORBmatcher.cc:1332-1474; scope row N2)
//   SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (replaces S/ORBmatcher.cc:1476-1603; scope row N2)
//   SearchByBoW(pKF, F, vpMapPointMatches)               (replaces S/ORBmatcher.cc:161-292; scope row N3)
//   SearchByBoW(pKF1, pKF2, vpMatches12)                 (replaces S/ORBmatcher.cc:526-659; scope row N3)
//   SearchForTriangulation(pKF1, pKF2, F12, pairs, bOnlyStereo) (replaces S/ORBmatcher.cc:661-827; scope row N3)
//   Fuse(pKF, vpMapPoints, th)                           (replaces S/ORBmatcher.cc:829-975; scope row N3: the search
//                                                         runs on the device, the replace-or-add surgery stays here)
//   Fuse(pKF, Scw, vpPoints, th, vpReplacePoint)         (replaces S/ORBmatcher.cc:979-1104; same split)
//   SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th) (replaces S/ORBmatcher.cc:1106-1330)
//   SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) (replaces S/ORBmatcher.cc:294-407)
// It compiles against the reference's own, unmodified headers (ORBmatcher.h, Frame.h, MapPoint.h),
// so Tracking.cc, LocalMapping.cc and LoopClosing.cc call them unchanged: build ORBmatcher.cc with
// -DORB_B200_MATCHER and guard the replaced bodies with #ifndef ORB_B200_MATCHER.
// The pointer graph is flattened to structure-of-arrays on the host, the search runs in CUDA
// through include/orb_b200.h, and the results are written back into the fields the reference
// mutates.  No CPU search path exists here; on a device error the call logs and reports 0 matches.
#include <atomic>
#include "ORBmatcher.h"

#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <vector>

#include "orb_b200.h"

namespace ORB_SLAM2
{

namespace
{
// One device scratch handle per host thread: ORBmatcher is stateless and is used concurrently from
// the Tracking, LocalMapping and LoopClosing threads (S/System.cc:156,160).
// The CUDA device the matcher handles of this process live on (orbb200_shim_set_matcher_device; default 0).  A thread's
// handle follows the setting the next time it is used.
std::atomic<int> gMatcherDevice(0);

struct ThreadMatcher {
    orbb200_matcher* h; int items, points, device;
    ThreadMatcher() : h(0), items(0), points(0), device(0) {}
    ~ThreadMatcher() { if (h) orbb200_matcher_destroy(h); }
    orbb200_matcher* get(int needPoints)
    {
        const int want = gMatcherDevice.load();
        if (h && needPoints <= points && device == want) return h;
        if (h) orbb200_matcher_destroy(h);
        h = 0;
        if (needPoints < points) needPoints = points;
        points = needPoints < 4096 ? 4096 : needPoints;
        device = want;
        if (orbb200_matcher_create(1, points, device, &h) != ORBB200_OK) {
            std::fprintf(stderr, "ORBmatcher(B200): %s\n", orbb200_last_error());
            h = 0; points = 0;
        }
        return h;
    }
};
thread_local ThreadMatcher tlsMatcher;

struct FrameSoA {
    int32_t n;
    std::vector<float> x, y, angle;
    std::vector<int32_t> octave;
    std::vector<unsigned char> desc;
    orbb200_frame_view view;
    explicit FrameSoA(const Frame& F)
    {
        n = (int32_t)F.mvKeysUn.size();
        const int s = n > 0 ? n : 1;
        x.resize(s); y.resize(s); angle.resize(s); octave.resize(s); desc.resize((size_t)s * 32);
        for (int i = 0; i < n; i++) {
            const cv::KeyPoint& kp = F.mvKeysUn[i];
            x[i] = kp.pt.x; y[i] = kp.pt.y; angle[i] = kp.angle; octave[i] = kp.octave;
            std::memcpy(&desc[(size_t)i * 32], F.mDescriptors.ptr(i), 32);
        }
        view.n = &n; view.x = &x[0]; view.y = &y[0]; view.octave = &octave[0]; view.angle = &angle[0];
        view.desc = &desc[0]; view.stride = s;
    }
};

void FrameBounds(float b[4])
{
    b[0] = Frame::mnMinX; b[1] = Frame::mnMinY; b[2] = Frame::mnMaxX; b[3] = Frame::mnMaxY;
}
}  // namespace

// Selects the CUDA device of the matcher handles (one per host thread); call it before the threads that match start, e.g.
// next to the ORBextractor's device argument.  Declared in the integrator's code as
//     namespace ORB_SLAM2 { void orbb200_shim_set_matcher_device(int device); }
void orbb200_shim_set_matcher_device(int device) { gMatcherDevice.store(device < 0 ? 0 : device); }

int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b)
{
    orbb200_matcher* h = tlsMatcher.get(1);
    int32_t dist = 0;
    if (!h || orbb200_descriptor_distance(h, a.ptr<unsigned char>(), b.ptr<unsigned char>(), 1, &dist) != ORBB200_OK) {
        std::fprintf(stderr, "ORBmatcher(B200)::DescriptorDistance: %s\n", orbb200_last_error());
        return 256;
    }
    return dist;
}

int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched,
                                        std::vector<int>& vnMatches12, int windowSize)
{
    vnMatches12 = std::vector<int>(F1.mvKeysUn.size(), -1);
    FrameSoA s1(F1), s2(F2);
    orbb200_matcher* h = tlsMatcher.get(s1.n > s2.n ? s1.n : s2.n);
    if (!h) return 0;
    std::vector<float> prev((size_t)s1.view.stride * 2, 0.f);
    for (int i = 0; i < s1.n; i++) { prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y; }
    std::vector<int32_t> m12(s1.view.stride, -1);
    int32_t nmatches = 0;
    float bounds[4];
    FrameBounds(bounds);
    if (orbb200_search_for_initialization(h, 1, &s1.view, &s2.view, bounds, mfNNratio, mbCheckOrientation ? 1 : 0,
                                          windowSize, &prev[0], &m12[0], &nmatches, 0) != ORBB200_OK) {
        std::fprintf(stderr, "ORBmatcher(B200)::SearchForInitialization: %s\n", orbb200_last_error());
        return 0;
    }
    for (int i = 0; i < s1.n; i++) {
        vnMatches12[i] = m12[i];
        vbPrevMatched[i] = cv::Point2f(prev[2 * i], prev[2 * i + 1]);
    }
    return nmatches;
}

int ORBmatcher::SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th)
{
    const int nmp = (int)vpMapPoints.size();
    FrameSoA s(F);
    orbb200_matcher* h = tlsMatcher.get(s.n > nmp ? s.n : nmp);
    if (!h || nmp == 0) return 0;

    // map points -> SoA (the "variables used by the tracking", I/MapPoint.h:96-104)
    const int ms = nmp;
    int32_t mn = nmp;
    std::vector<unsigned char> inView(ms), bad(ms), desc((size_t)ms * 32);
    std::vector<float> px(ms), py(ms), pxr(ms), vcos(ms);
    std::vector<int32_t> level(ms), obs(ms);
    std::map<const MapPoint*, int> indexOf;
    for (int i = 0; i < nmp; i++) {
        MapPoint* pMP = vpMapPoints[i];
        indexOf[pMP] = i;
        inView[i] = pMP->mbTrackInView ? 1 : 0;
        bad[i] = pMP->isBad() ? 1 : 0;
        px[i] = pMP->mTrackProjX; py[i] = pMP->mTrackProjY; pxr[i] = pMP->mTrackProjXR;
        level[i] = pMP->mnTrackScaleLevel; vcos[i] = pMP->mTrackViewCos;
        obs[i] = pMP->Observations();
        const cv::Mat d = pMP->GetDescriptor();
        if (!d.empty()) std::memcpy(&desc[(size_t)i * 32], d.ptr<unsigned char>(), 32);
    }
    orbb200_mappoint_view mv;
    mv.n = &mn; mv.in_view = &inView[0]; mv.bad = &bad[0]; mv.proj_x = &px[0]; mv.proj_y = &py[0]; mv.proj_xr = &pxr[0];
    mv.level = &level[0]; mv.view_cos = &vcos[0]; mv.desc = &desc[0]; mv.obs = &obs[0]; mv.stride = ms;

    // keypoint occupancy (Frame::mvpMapPoints): index into vpMapPoints, or -2 + Observations() for others
    std::vector<int32_t> kpMp(s.view.stride, -1), kpObs(s.view.stride, 0);
    std::vector<float> uRight(s.view.stride, -1.f);
    for (int i = 0; i < s.n; i++) {
        uRight[i] = F.mvuRight[i];
        MapPoint* held = F.mvpMapPoints[i];
        if (!held) continue;
        std::map<const MapPoint*, int>::const_iterator it = indexOf.find(held);
        if (it != indexOf.end()) kpMp[i] = it->second;
        else { kpMp[i] = -2; kpObs[i] = held->Observations(); }
    }
    const std::vector<int32_t> before(kpMp);

    float bounds[4];
    FrameBounds(bounds);
    int32_t nmatches = 0;
    if (orbb200_search_by_projection(h, 1, &s.view, &uRight[0], &mv, &kpMp[0], &kpObs[0], &F.mvScaleFactors[0],
                                     (int)F.mvScaleFactors.size(), bounds, mfNNratio, th, &nmatches, 0) != ORBB200_OK) {
        std::fprintf(stderr, "ORBmatcher(B200)::SearchByProjection: %s\n", orbb200_last_error());
        return 0;
    }
    for (int i = 0; i < s.n; i++)
        if (kpMp[i] != before[i] && kpMp[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[kpMp[i]];   // :125
    return nmatches;
}

int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono)
{
    // the two scalars that pick the octave window are evaluated on the host exactly as in the reference (:1342-1353)
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat twc = -Rcw.t() * tcw;
    const cv::Mat Rlw = LastFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tlw = LastFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat tlc = Rlw * twc + tlw;
    const bool bForward = tlc.at<float>(2) > CurrentFrame.mb && !bMono;
    const bool bBackward = -tlc.at<float>(2) > CurrentFrame.mb && !bMono;
    const int mode = bForward ? 1 : (bBackward ? 2 : 0);

    FrameSoA cur(CurrentFrame);
    const int nl = LastFrame.N;
    orbb200_matcher* h = tlsMatcher.get(cur.n > nl ? cur.n : nl);
    if (!h || nl == 0) return 0;

    const int ls = nl;
    int32_t ln = nl;
    std::vector<unsigned char> hasMp(ls), outlier(ls), desc((size_t)ls * 32);
    std::vector<float> wpos((size_t)ls * 3), angle(ls);
    std::vector<int32_t> obs(ls), octave(ls);
    std::map<const MapPoint*, int> indexOf;
    for (int i = 0; i < nl; i++) {
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        hasMp[i] = pMP ? 1 : 0;
        outlier[i] = LastFrame.mvbOutlier[i] ? 1 : 0;
        octave[i] = LastFrame.mvKeys[i].octave;
        angle[i] = LastFrame.mvKeysUn[i].angle;
        if (!pMP) continue;
        indexOf[pMP] = i;
        const cv::Mat x3Dw = pMP->GetWorldPos();
        for (int k = 0; k < 3; k++) wpos[3 * (size_t)i + k] = x3Dw.at<float>(k);
        const cv::Mat d = pMP->GetDescriptor();
        if (!d.empty()) std::memcpy(&desc[(size_t)i * 32], d.ptr<unsigned char>(), 32);
        obs[i] = pMP->Observations();
    }
    orbb200_lastframe_view lv;
    lv.n = &ln; lv.has_mp = &hasMp[0]; lv.outlier = &outlier[0]; lv.world_pos = &wpos[0]; lv.mp_desc = &desc[0];
    lv.mp_obs = &obs[0]; lv.octave = &octave[0]; lv.angle = &angle[0]; lv.stride = ls;

    std::vector<int32_t> kpMp(cur.view.stride, -1), kpObs(cur.view.stride, 0);
    std::vector<float> uRight(cur.view.stride, -1.f);
    for (int i = 0; i < cur.n; i++) {
        uRight[i] = CurrentFrame.mvuRight[i];
        MapPoint* held = CurrentFrame.mvpMapPoints[i];
        if (!held) continue;
        std::map<const MapPoint*, int>::const_iterator it = indexOf.find(held);
        if (it != indexOf.end()) kpMp[i] = it->second;
        else { kpMp[i] = -2; kpObs[i] = held->Observations(); }
    }
    const std::vector<int32_t> before(kpMp);
    float R9[9], t3[3];
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) R9[3 * r + c] = Rcw.at<float>(r, c); t3[r] = tcw.at<float>(r); }
    const float K[4] = {CurrentFrame.fx, CurrentFrame.fy, CurrentFrame.cx, CurrentFrame.cy};
    float bounds[4];
    FrameBounds(bounds);
    int32_t nmatches = 0;
    if (orbb200_search_by_projection_last_frame(h, 1, &cur.view, &uRight[0], &lv, R9, t3, K, CurrentFrame.mbf, &kpMp[0], &kpObs[0],
                                                &CurrentFrame.mvScaleFactors[0], (int)CurrentFrame.mvScaleFactors.size(), bounds,
                                                th, mode, mbCheckOrientation ? 1 : 0, &nmatches, 0) != ORBB200_OK) {
        std::fprintf(stderr, "ORBmatcher(B200)::SearchByProjection(last frame): %s\n", orbb200_last_error());
        return 0;
    }
    for (int i = 0; i < cur.n; i++) {
        if (kpMp[i] == before[i]) continue;
        CurrentFrame.mvpMapPoints[i] = kpMp[i] >= 0 ? LastFrame.mvpMapPoints[kpMp[i]] : static_cast<MapPoint*>(NULL);   // :1438, :1465
    }
    return nmatches;
}

namespace
{
// MapPoint keeps mfMaxDistance / mfMinDistance protected and only exposes them multiplied by 1.2f / 0.8f; the
// device needs the raw values (PredictScale uses the raw maximum).  A derived class may name the protected members
// of its base, and the resulting pointers to members have type `float MapPoint::*`: no object of this type exists.
struct MapPointFields : public MapPoint {
    static float MapPoint::* MaxDistance() { return &MapPointFields::mfMaxDistance; }
    static float MapPoint::* MinDistance() { return &MapPointFields::mfMinDistance; }
    static std::mutex MapPoint::* PosMutex() { return &MapPointFields::mMutexPos; }
};
}  // namespace

int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th,
                                   const int ORBdist)
{
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat Ow = -Rcw.t() * tcw;                                                 // :1482, evaluated by OpenCV as before
    const std::vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();

    FrameSoA cur(CurrentFrame);
    const int nk = (int)vpMPs.size();
    orbb200_matcher* h = tlsMatcher.get(cur.n > nk ? cur.n : nk);
    if (!h || nk == 0) return 0;

    int32_t kn = nk;
    std::vector<unsigned char> valid(nk), desc((size_t)nk * 32);
    std::vector<float> wpos((size_t)nk * 3), angle(nk), maxD(nk), minD(nk);
    for (int i = 0; i < nk; i++) {
        MapPoint* pMP = vpMPs[i];
        angle[i] = pKF->mvKeysUn[i].angle;
        valid[i] = (pMP && !pMP->isBad() && !sAlreadyFound.count(pMP)) ? 1 : 0;        // :1497-1499
        if (!valid[i]) continue;
        const cv::Mat x3Dw = pMP->GetWorldPos();
        for (int k = 0; k < 3; k++) wpos[3 * (size_t)i + k] = x3Dw.at<float>(k);
        const cv::Mat d = pMP->GetDescriptor();
        if (!d.empty()) std::memcpy(&desc[(size_t)i * 32], d.ptr<unsigned char>(), 32);
        std::unique_lock<std::mutex> lock(pMP->*MapPointFields::PosMutex());
        maxD[i] = pMP->*MapPointFields::MaxDistance();
        minD[i] = pMP->*MapPointFields::MinDistance();
    }
    orbb200_keyframe_view kv;
    kv.n = &kn; kv.valid = &valid[0]; kv.world_pos = &wpos[0]; kv.mp_desc = &desc[0]; kv.max_distance = &maxD[0];
    kv.min_distance = &minD[0]; kv.angle = &angle[0]; kv.stride = nk;

    std::vector<int32_t> kpMp(cur.view.stride, -1);
    for (int i = 0; i < cur.n; i++)
        if (CurrentFrame.mvpMapPoints[i]) kpMp[i] = -2;                                // any held map point blocks the keypoint (:1546)
    const std::vector<int32_t> before(kpMp);
    float R9[9], t3[3], O3[3];
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) R9[3 * r + c] = Rcw.at<float>(r, c); t3[r] = tcw.at<float>(r); O3[r] = Ow.at<float>(r); }
    const float K[4] = {CurrentFrame.fx, CurrentFrame.fy, CurrentFrame.cx, CurrentFrame.cy};
    float bounds[4];
    FrameBounds(bounds);
    int32_t nmatches = 0;
    if (orbb200_search_by_projection_keyframe(h, 1, &cur.view, &kv, R9, t3, O3, K, &kpMp[0], &CurrentFrame.mvScaleFactors[0],
                                              (int)CurrentFrame.mvScaleFactors.size(), CurrentFrame.mfLogScaleFactor, bounds, th, ORBdist,
                                              mbCheckOrientation ? 1 : 0, &nmatches, 0) != ORBB200_OK) {
        std::fprintf(stderr, "ORBmatcher(B200)::SearchByProjection(key frame): %s\n", orbb200_last_error());
        return 0;
    }
    for (int i = 0; i < cur.n; i++)
        if (kpMp[i] != before[i] && kpMp[i] >= 0) CurrentFrame.mvpMapPoints[i] = vpMPs[kpMp[i]];   // :1563 (rejected ones stay NULL, :1594)
    return nmatches;
}

namespace
{
// DBoW2::FeatureVector -> the flat arrays of orbb200_bow_view (node ids in map order, offsets, feature indices)
struct FlatFeatVec {
    int32_t nNodes;
    std::vector<uint32_t> node, feat;
    std::vector<int32_t> start;
    FlatFeatVec(const DBoW2::FeatureVector& fv, int nFeatures)
    {
        start.push_back(0);
        for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
            node.push_back(it->first);
            feat.insert(feat.end(), it->second.begin(), it->second.end());
            start.push_back((int32_t)feat.size());
        }
        nNodes = (int32_t)node.size();
        if (node.empty()) node.push_back(0);
        feat.resize(nFeatures > (int)feat.size() ? nFeatures : feat.size(), 0);
        if (feat.empty()) feat.push_back(0);
    }
};
}  // namespace

int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches)
{
    const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = std::vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
    const int nk = (int)vpMapPointsKF.size(), nf = F.N;
    if (nk == 0 || nf == 0) return 0;
    orbb200_matcher* h = tlsMatcher.get(nk > nf ? nk : nf);
    if (!h) return 0;

    int32_t kn = nk, fn = nf;
    std::vector<unsigned char> valid(nk), kdesc((size_t)nk * 32), fdesc((size_t)nf * 32);
    std::vector<float> kang(nk), fang(nf);
    for (int i = 0; i < nk; i++) {
        MapPoint* pMP = vpMapPointsKF[i];
        valid[i] = (pMP && !pMP->isBad()) ? 1 : 0;                                     // :193-198
        kang[i] = pKF->mvKeysUn[i].angle;
        std::memcpy(&kdesc[(size_t)i * 32], pKF->mDescriptors.ptr<unsigned char>(i), 32);
    }
    for (int i = 0; i < nf; i++) {
        fang[i] = F.mvKeys[i].angle;                                                   // (the reference reads mvKeys here, :240)
        std::memcpy(&fdesc[(size_t)i * 32], F.mDescriptors.ptr<unsigned char>(i), 32);
    }
    FlatFeatVec kfv(pKF->mFeatVec, nk), ffv(F.mFeatVec, nf);
    orbb200_bow_view kv, fv;
    kv.n = &kn; kv.desc = &kdesc[0]; kv.angle = &kang[0]; kv.valid = &valid[0]; kv.n_nodes = &kfv.nNodes; kv.node_id = &kfv.node[0];
    kv.node_start = &kfv.start[0]; kv.feat = &kfv.feat[0]; kv.stride = nk; kv.node_stride = (int)kfv.node.size();
    fv.n = &fn; fv.desc = &fdesc[0]; fv.angle = &fang[0]; fv.valid = 0; fv.n_nodes = &ffv.nNodes; fv.node_id = &ffv.node[0];
    fv.node_start = &ffv.start[0]; fv.feat = &ffv.feat[0]; fv.stride = nf; fv.node_stride = (int)ffv.node.size();
    kfv.start.resize(kv.node_stride + 1, kfv.start.back());
    ffv.start.resize(fv.node_stride + 1, ffv.start.back());
    kv.node_start = &kfv.start[0]; fv.node_start = &ffv.start[0];

    std::vector<int32_t> matches(nf, -1);
    int32_t nmatches = 0;
    if (orbb200_search_by_bow(h, 1, &kv, &fv, mfNNratio, mbCheckOrientation ? 1 : 0, &matches[0], &nmatches, 0) != ORBB200_OK) {
        std::fprintf(stderr, "ORBmatcher(B200)::SearchByBoW: %s\n", orbb200_last_error());
        return 0;
    }
    for (int i = 0; i < nf; i++)
        if (matches[i] >= 0) vpMapPointMatches[i] = vpMapPointsKF[matches[i]];         // :234
    return nmatches;
}

namespace
{
// one key frame as an orbb200_bow_view (buffers owned by the object)
struct KeyFrameBow {
    int32_t n;
    std::vector<MapPoint*> mps;
    std::vector<unsigned char> valid, desc;
    std::vector<float> angle;
    FlatFeatVec fv;
    orbb200_bow_view view;
    explicit KeyFrameBow(KeyFrame* pKF) : mps(pKF->GetMapPointMatches()), fv(pKF->mFeatVec, (int)mps.size())
    {
        n = (int32_t)mps.size();
        const int s = n > 0 ? n : 1;
        valid.resize(s); desc.resize((size_t)s * 32); angle.resize(s);
        for (int i = 0; i < n; i++) {
            valid[i] = (mps[i] && !mps[i]->isBad()) ? 1 : 0;
            angle[i] = pKF->mvKeysUn[i].angle;
            std::memcpy(&desc[(size_t)i * 32], pKF->mDescriptors.ptr<unsigned char>(i), 32);
        }
        fv.start.resize(fv.node.size() + 1, fv.start.back());
        view.n = &n; view.desc = &desc[0]; view.angle = &angle[0]; view.valid = &valid[0]; view.n_nodes = &fv.nNodes;
        view.node_id = &fv.node[0]; view.node_start = &fv.start[0]; view.feat = &fv.feat[0]; view.stride = s;
        view.node_stride = (int)fv.node.size();
    }
};
}  // namespace

int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12)
{
    KeyFrameBow k1(pKF1), k2(pKF2);
    vpMatches12 = std::vector<MapPoint*>(k1.mps.size(), static_cast<MapPoint*>(NULL));
    if (k1.n == 0 || k2.n == 0) return 0;
    orbb200_matcher* h = tlsMatcher.get(k1.n > k2.n ? k1.n : k2.n);
    if (!h) return 0;
    std::vector<int32_t> matches(k1.n, -1);
    int32_t nmatches = 0;
    if (orbb200_search_by_bow_keyframes(h, 1, &k1.view, &k2.view, mfNNratio, mbCheckOrientation ? 1 : 0, &matches[0], &nmatches, 0) != ORBB200_OK) {
        std::fprintf(stderr, "ORBmatcher(B200)::SearchByBoW(KF, KF): %s\n", orbb200_last_error());
        return 0;
    }
    for (int i = 0; i < k1.n; i++)
        if (matches[i] >= 0) vpMatches12[i] = k2.mps[matches[i]];                      // :605
    return nmatches;
}

namespace
{
struct KeyFrameGeo {
    std::vector<float> x, y, uRight;
    std::vector<int32_t> octave;
    std::vector<unsigned char> hasMp;
    orbb200_tri_view view;
    KeyFrameGeo(KeyFrame* pKF, const KeyFrameBow& b)
    {
        const int s = b.view.stride;
        x.resize(s); y.resize(s); uRight.resize(s, -1.f); octave.resize(s); hasMp.resize(s);
        for (int i = 0; i < b.n; i++) {
            x[i] = pKF->mvKeysUn[i].pt.x; y[i] = pKF->mvKeysUn[i].pt.y; octave[i] = pKF->mvKeysUn[i].octave;
            uRight[i] = pKF->mvuRight[i];
            hasMp[i] = b.mps[i] ? 1 : 0;                                               // :706, :726 (a bad map point blocks too)
        }
        view.x = &x[0]; view.y = &y[0]; view.octave = &octave[0]; view.u_right = &uRight[0]; view.has_mp = &hasMp[0];
    }
};
}  // namespace

int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs,
                                       const bool bOnlyStereo)
{
    // epipole in the second image, evaluated on the host exactly as in the reference (:668-675)
    cv::Mat Cw = pKF1->GetCameraCenter();
    cv::Mat R2w = pKF2->GetRotation();
    cv::Mat t2w = pKF2->GetTranslation();
    cv::Mat C2 = R2w * Cw + t2w;
    const float invz = 1.0f / C2.at<float>(2);
    const float epipole[2] = {pKF2->fx * C2.at<float>(0) * invz + pKF2->cx, pKF2->fy * C2.at<float>(1) * invz + pKF2->cy};

    vMatchedPairs.clear();
    KeyFrameBow k1(pKF1), k2(pKF2);
    if (k1.n == 0 || k2.n == 0) return 0;
    KeyFrameGeo g1(pKF1, k1), g2(pKF2, k2);
    orbb200_matcher* h = tlsMatcher.get(k1.n > k2.n ? k1.n : k2.n);
    if (!h) return 0;
    float F9[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) F9[3 * r + c] = F12.at<float>(r, c);
    std::vector<int32_t> matches(k1.n, -1);
    int32_t nmatches = 0;
    if (orbb200_search_for_triangulation(h, 1, &k1.view, &g1.view, &k2.view, &g2.view, F9, epipole, &pKF2->mvScaleFactors[0],
                                         &pKF2->mvLevelSigma2[0], (int)pKF2->mvScaleFactors.size(), bOnlyStereo ? 1 : 0,
                                         mbCheckOrientation ? 1 : 0, &matches[0], &nmatches, 0) != ORBB200_OK) {
        std::fprintf(stderr, "ORBmatcher(B200)::SearchForTriangulation: %s\n", orbb200_last_error());
        return 0;
    }
    vMatchedPairs.reserve(nmatches);
    for (int i = 0; i < k1.n; i++)
        if (matches[i] >= 0) vMatchedPairs.push_back(std::make_pair((size_t)i, (size_t)matches[i]));      // :815-820
    return nmatches;
}

int ORBmatcher::Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th)
{
    cv::Mat Rcw = pKF->GetRotation();
    cv::Mat tcw = pKF->GetTranslation();
    cv::Mat Ow = pKF->GetCameraCenter();
    const int nMPs = (int)vpMapPoints.size(), nk = pKF->N;
    if (nMPs == 0 || nk == 0) return 0;
    orbb200_matcher* h = tlsMatcher.get(nMPs > nk ? nMPs : nk);
    if (!h) return 0;

    // the key frame's keypoints
    int32_t kn = nk;
    std::vector<float> kx(nk), ky(nk), kur(nk);
    std::vector<int32_t> koct(nk);
    std::vector<unsigned char> kdesc((size_t)nk * 32);
    for (int i = 0; i < nk; i++) {
        kx[i] = pKF->mvKeysUn[i].pt.x; ky[i] = pKF->mvKeysUn[i].pt.y; koct[i] = pKF->mvKeysUn[i].octave; kur[i] = pKF->mvuRight[i];
        std::memcpy(&kdesc[(size_t)i * 32], pKF->mDescriptors.ptr<unsigned char>(i), 32);
    }
    orbb200_frame_view kv;
    kv.n = &kn; kv.x = &kx[0]; kv.y = &ky[0]; kv.octave = &koct[0]; kv.angle = 0; kv.desc = &kdesc[0]; kv.stride = nk;

    // the candidates, in the state the reference would find them in when it reaches them (see below)
    int32_t pn = nMPs;
    std::vector<unsigned char> valid(nMPs), desc((size_t)nMPs * 32);
    std::vector<float> wpos((size_t)nMPs * 3), normal((size_t)nMPs * 3), maxD(nMPs), minD(nMPs);
    for (int i = 0; i < nMPs; i++) {
        MapPoint* pMP = vpMapPoints[i];
        valid[i] = (pMP && !pMP->isBad() && !pMP->IsInKeyFrame(pKF)) ? 1 : 0;          // :845-849
        if (!valid[i]) continue;
        const cv::Mat p3Dw = pMP->GetWorldPos(), Pn = pMP->GetNormal(), d = pMP->GetDescriptor();
        for (int k = 0; k < 3; k++) { wpos[3 * (size_t)i + k] = p3Dw.at<float>(k); normal[3 * (size_t)i + k] = Pn.at<float>(k); }
        if (!d.empty()) std::memcpy(&desc[(size_t)i * 32], d.ptr<unsigned char>(), 32);
        std::unique_lock<std::mutex> lock(pMP->*MapPointFields::PosMutex());
        maxD[i] = pMP->*MapPointFields::MaxDistance();
        minD[i] = pMP->*MapPointFields::MinDistance();
    }
    orbb200_fusepoints_view pv;
    pv.n = &pn; pv.valid = &valid[0]; pv.world_pos = &wpos[0]; pv.normal = &normal[0]; pv.mp_desc = &desc[0];
    pv.max_distance = &maxD[0]; pv.min_distance = &minD[0]; pv.stride = nMPs;

    float R9[9], t3[3], O3[3];
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) R9[3 * r + c] = Rcw.at<float>(r, c); t3[r] = tcw.at<float>(r); O3[r] = Ow.at<float>(r); }
    const float K[4] = {pKF->fx, pKF->fy, pKF->cx, pKF->cy};
    float bounds[4];
    FrameBounds(bounds);                     // the Frame statics the key frame's grid was assigned with
    std::vector<int32_t> best(nMPs, -1);
    if (orbb200_fuse_search(h, 1, &kv, &kur[0], &pv, R9, t3, O3, K, pKF->mbf, &pKF->mvScaleFactors[0], &pKF->mvInvLevelSigma2[0],
                            (int)pKF->mvScaleFactors.size(), pKF->mfLogScaleFactor, bounds, th, 0, 0, 0, &best[0], 0, 0) != ORBB200_OK) {
        std::fprintf(stderr, "ORBmatcher(B200)::Fuse: %s\n", orbb200_last_error());
        return 0;
    }

    // The surgery, in list order, exactly as :950-971.  The search result of a candidate does not depend on earlier
    // surgery (keypoints are not consumed, positions and descriptors of unvisited candidates are untouched); its
    // admission does: a candidate that earlier surgery made bad, or attached to this key frame, is skipped here as
    // the reference would skip it (a candidate cannot become admissible again).
    int nFused = 0;
    for (int i = 0; i < nMPs; i++) {
        MapPoint* pMP = vpMapPoints[i];
        if (!pMP || !valid[i] || best[i] < 0) continue;
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
        const int bestIdx = best[i];
        MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) {
                if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                else pMPinKF->Replace(pMP);
            }
        } else {
            pMP->AddObservation(pKF, bestIdx);
            pKF->AddMapPoint(pMP, bestIdx);
        }
        nFused++;
    }
    return nFused;
}

namespace
{
// a key frame's keypoints as an orbb200_frame_view (+ mvuRight)
struct KeyFrameKeys {
    int32_t n;
    std::vector<float> x, y, uRight;
    std::vector<int32_t> octave;
    std::vector<unsigned char> desc;
    orbb200_frame_view view;
    explicit KeyFrameKeys(KeyFrame* pKF)
    {
        n = pKF->N;
        const int s = n > 0 ? n : 1;
        x.resize(s); y.resize(s); uRight.resize(s, -1.f); octave.resize(s); desc.resize((size_t)s * 32);
        for (int i = 0; i < n; i++) {
            x[i] = pKF->mvKeysUn[i].pt.x; y[i] = pKF->mvKeysUn[i].pt.y; octave[i] = pKF->mvKeysUn[i].octave; uRight[i] = pKF->mvuRight[i];
            std::memcpy(&desc[(size_t)i * 32], pKF->mDescriptors.ptr<unsigned char>(i), 32);
        }
        view.n = &n; view.x = &x[0]; view.y = &y[0]; view.octave = &octave[0]; view.angle = 0; view.desc = &desc[0]; view.stride = s;
    }
};

// candidate map points as an orbb200_fusepoints_view; admit[i] decides which are looked at
struct PointSoA {
    int32_t n;
    std::vector<unsigned char> valid, desc;
    std::vector<float> wpos, normal, maxD, minD;
    orbb200_fusepoints_view view;
    PointSoA(const std::vector<MapPoint*>& pts, const std::vector<unsigned char>& admit)
    {
        n = (int32_t)pts.size();
        const int s = n > 0 ? n : 1;
        valid.assign(s, 0); desc.resize((size_t)s * 32); wpos.resize((size_t)s * 3); normal.resize((size_t)s * 3); maxD.resize(s); minD.resize(s);
        for (int i = 0; i < n; i++) {
            MapPoint* pMP = pts[i];
            if (!admit[i]) continue;
            valid[i] = 1;
            const cv::Mat p3Dw = pMP->GetWorldPos(), Pn = pMP->GetNormal(), d = pMP->GetDescriptor();
            for (int k = 0; k < 3; k++) { wpos[3 * (size_t)i + k] = p3Dw.at<float>(k); normal[3 * (size_t)i + k] = Pn.at<float>(k); }
            if (!d.empty()) std::memcpy(&desc[(size_t)i * 32], d.ptr<unsigned char>(), 32);
            std::unique_lock<std::mutex> lock(pMP->*MapPointFields::PosMutex());
            maxD[i] = pMP->*MapPointFields::MaxDistance();
            minD[i] = pMP->*MapPointFields::MinDistance();
        }
        view.n = &n; view.valid = &valid[0]; view.world_pos = &wpos[0]; view.normal = &normal[0]; view.mp_desc = &desc[0];
        view.max_distance = &maxD[0]; view.min_distance = &minD[0]; view.stride = s;
    }
};

void Flatten3x3(const cv::Mat& R, float* out9) { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) out9[3 * r + c] = R.at<float>(r, c); }
void Flatten3(const cv::Mat& t, float* out3) { for (int r = 0; r < 3; r++) out3[r] = t.at<float>(r); }
}  // namespace

int ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint)
{
    // Decompose Scw on the host, as the reference does (:987-991)
    cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    cv::Mat Rcw = sRcw / scw;
    cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
    cv::Mat Ow = -Rcw.t() * tcw;
    const std::set<MapPoint*> spAlreadyFound = pKF->GetMapPoints();

    const int nPoints = (int)vpPoints.size();
    if (nPoints == 0 || pKF->N == 0) return 0;
    orbb200_matcher* h = tlsMatcher.get(nPoints > pKF->N ? nPoints : pKF->N);
    if (!h) return 0;
    std::vector<unsigned char> admit(nPoints);
    for (int i = 0; i < nPoints; i++) admit[i] = (!vpPoints[i]->isBad() && !spAlreadyFound.count(vpPoints[i])) ? 1 : 0;   // :1005-1006
    KeyFrameKeys keys(pKF);
    PointSoA pts(vpPoints, admit);
    float R9[9], t3[3], O3[3], bounds[4];
    Flatten3x3(Rcw, R9); Flatten3(tcw, t3); Flatten3(Ow, O3);
    FrameBounds(bounds);
    const float K[4] = {pKF->fx, pKF->fy, pKF->cx, pKF->cy};
    std::vector<int32_t> best(nPoints, -1);
    if (orbb200_fuse_search(h, 1, &keys.view, &keys.uRight[0], &pts.view, R9, t3, O3, K, pKF->mbf, &pKF->mvScaleFactors[0],
                            &pKF->mvInvLevelSigma2[0], (int)pKF->mvScaleFactors.size(), pKF->mfLogScaleFactor, bounds, th, 1, 0, 0,
                            &best[0], 0, 0) != ORBB200_OK) {
        std::fprintf(stderr, "ORBmatcher(B200)::Fuse(Scw): %s\n", orbb200_last_error());
        return 0;
    }
    int nFused = 0;
    for (int iMP = 0; iMP < nPoints; iMP++) {                                          // :1084-1100, in list order
        MapPoint* pMP = vpPoints[iMP];
        if (!admit[iMP] || best[iMP] < 0 || pMP->isBad()) continue;
        MapPoint* pMPinKF = pKF->GetMapPoint(best[iMP]);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) vpReplacePoint[iMP] = pMPinKF;
        } else {
            pMP->AddObservation(pKF, best[iMP]);
            pKF->AddMapPoint(pMP, best[iMP]);
        }
        nFused++;
    }
    return nFused;
}

int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12,
                             const cv::Mat& t12, const float th)
{
    // the transforms, evaluated on the host as in the reference (:1114-1126)
    cv::Mat R1w = pKF1->GetRotation(), t1w = pKF1->GetTranslation(), R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
    cv::Mat sR12 = s12 * R12;
    cv::Mat sR21 = (1.0 / s12) * R12.t();
    cv::Mat t21 = -sR21 * t12;

    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
    const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
    if (N1 == 0 || N2 == 0) return 0;
    std::vector<unsigned char> admit1(N1, 0), admit2(N2, 0);
    std::vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
    for (int i = 0; i < N1; i++) {                                                     // :1137-1148
        MapPoint* pMP = vpMatches12[i];
        if (pMP) {
            vbAlreadyMatched1[i] = true;
            const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
            if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
        }
    }
    for (int i = 0; i < N1; i++) admit1[i] = (vpMapPoints1[i] && !vbAlreadyMatched1[i] && !vpMapPoints1[i]->isBad()) ? 1 : 0;
    for (int i = 0; i < N2; i++) admit2[i] = (vpMapPoints2[i] && !vbAlreadyMatched2[i] && !vpMapPoints2[i]->isBad()) ? 1 : 0;

    orbb200_matcher* h = tlsMatcher.get(N1 > N2 ? N1 : N2);
    if (!h) return 0;
    KeyFrameKeys keys1(pKF1), keys2(pKF2);
    PointSoA pts1(vpMapPoints1, admit1), pts2(vpMapPoints2, admit2);
    float bounds[4];
    FrameBounds(bounds);
    std::vector<int32_t> vnMatch1(N1, -1), vnMatch2(N2, -1);
    float Ra[9], ta[3], Rb[9], tb[3];
    // key frame 1's map points into key frame 2 (:1154-1230): camera 1, then sR21 / t21
    Flatten3x3(R1w, Ra); Flatten3(t1w, ta); Flatten3x3(sR21, Rb); Flatten3(t21, tb);
    const float K2[4] = {pKF1->fx, pKF1->fy, pKF1->cx, pKF1->cy};                      // the reference projects both legs with pKF1's intrinsics (:1109-1112)
    int rc = orbb200_fuse_search(h, 1, &keys2.view, 0, &pts1.view, Ra, ta, 0, K2, 0.f, &pKF2->mvScaleFactors[0], &pKF2->mvInvLevelSigma2[0],
                                 (int)pKF2->mvScaleFactors.size(), pKF2->mfLogScaleFactor, bounds, th, 2, Rb, tb, &vnMatch1[0], 0, 0);
    // key frame 2's map points into key frame 1 (:1233-1309): camera 2, then sR12 / t12
    Flatten3x3(R2w, Ra); Flatten3(t2w, ta); Flatten3x3(sR12, Rb); Flatten3(t12, tb);
    if (rc == ORBB200_OK)
        rc = orbb200_fuse_search(h, 1, &keys1.view, 0, &pts2.view, Ra, ta, 0, K2, 0.f, &pKF1->mvScaleFactors[0], &pKF1->mvInvLevelSigma2[0],
                                 (int)pKF1->mvScaleFactors.size(), pKF1->mfLogScaleFactor, bounds, th, 2, Rb, tb, &vnMatch2[0], 0, 0);
    if (rc != ORBB200_OK) {
        std::fprintf(stderr, "ORBmatcher(B200)::SearchBySim3: %s\n", orbb200_last_error());
        return 0;
    }
    int nFound = 0;                                                                    // agreement (:1312-1328)
    for (int i1 = 0; i1 < N1; i1++) {
        const int idx2 = vnMatch1[i1];
        if (idx2 >= 0 && vnMatch2[idx2] == i1) { vpMatches12[i1] = vpMapPoints2[idx2]; nFound++; }
    }
    return nFound;
}

int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th)
{
    cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);                                  // :303-308, on the host as before
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    cv::Mat Rcw = sRcw / scw;
    cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
    cv::Mat Ow = -Rcw.t() * tcw;
    std::set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
    spAlreadyFound.erase(static_cast<MapPoint*>(NULL));

    const int nPoints = (int)vpPoints.size(), nk = pKF->N;
    if (nPoints == 0 || nk == 0) return 0;
    orbb200_matcher* h = tlsMatcher.get(nPoints > nk ? nPoints : nk);
    if (!h) return 0;
    std::vector<unsigned char> admit(nPoints);
    for (int i = 0; i < nPoints; i++) admit[i] = (!vpPoints[i]->isBad() && !spAlreadyFound.count(vpPoints[i])) ? 1 : 0;   // :320-321
    KeyFrameKeys keys(pKF);
    PointSoA pts(vpPoints, admit);
    std::vector<int32_t> matched(keys.view.stride, -1);
    for (int i = 0; i < nk; i++) if (vpMatched[i]) matched[i] = -2;                    // occupied by an earlier match (:366)
    const std::vector<int32_t> before(matched);
    float R9[9], t3[3], O3[3], bounds[4];
    Flatten3x3(Rcw, R9); Flatten3(tcw, t3); Flatten3(Ow, O3);
    FrameBounds(bounds);
    const float K[4] = {pKF->fx, pKF->fy, pKF->cx, pKF->cy};
    int32_t nmatches = 0;
    if (orbb200_search_by_projection_sim3(h, 1, &keys.view, &pts.view, R9, t3, O3, K, &pKF->mvScaleFactors[0], (int)pKF->mvScaleFactors.size(),
                                          pKF->mfLogScaleFactor, bounds, th, &matched[0], &nmatches, 0) != ORBB200_OK) {
        std::fprintf(stderr, "ORBmatcher(B200)::SearchByProjection(Scw): %s\n", orbb200_last_error());
        return 0;
    }
    for (int i = 0; i < nk; i++)
        if (matched[i] != before[i] && matched[i] >= 0) vpMatched[i] = vpPoints[matched[i]];      // :385
    return nmatches;
}

}  // namespace ORB_SLAM2
