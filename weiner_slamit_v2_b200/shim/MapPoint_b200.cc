// MapPoint_b200.cc -- B200 body for ORB_SLAM2::MapPoint::ComputeDistinctiveDescriptors
// (replaces S/MapPoint.cc:248-313; scope row N4).  Guard the reference body with #ifndef ORB_B200_MAPPOINT and add
// this file.  The observed descriptors are gathered exactly as the reference gathers them (same
// locks, same std::map order, bad key frames skipped); the n x n Hamming medians run in CUDA through
// include/orb_b200.h.  LocalMapping calls this once per map point; a caller that holds many map points should use
// orbb200_distinctive_descriptors directly with all of them in one call.
#include "MapPoint.h"

#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "KeyFrame.h"
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
        if (!h && orbb200_matcher_create(1, 64, 0, &h) != ORBB200_OK) {
            std::fprintf(stderr, "MapPoint(B200): %s\n", orbb200_last_error());
            h = 0;
        }
        return h;
    }
};
thread_local ThreadHandle tlsHandle;
}  // namespace

void MapPoint::ComputeDistinctiveDescriptors()
{
    std::vector<cv::Mat> vDescriptors;
    std::map<KeyFrame*, size_t> observations;
    {
        std::unique_lock<std::mutex> lock1(mMutexFeatures);
        if (mbBad) return;
        observations = mObservations;
    }
    if (observations.empty()) return;
    vDescriptors.reserve(observations.size());
    for (std::map<KeyFrame*, size_t>::iterator mit = observations.begin(), mend = observations.end(); mit != mend; mit++) {
        KeyFrame* pKF = mit->first;
        if (!pKF->isBad()) vDescriptors.push_back(pKF->mDescriptors.row(mit->second));
    }
    if (vDescriptors.empty()) return;

    const int N = (int)vDescriptors.size();
    std::vector<unsigned char> flat((size_t)N * 32);
    for (int i = 0; i < N; i++) std::memcpy(&flat[(size_t)i * 32], vDescriptors[i].ptr<unsigned char>(), 32);
    const int32_t offsets[2] = {0, N};
    int32_t best = 0;
    orbb200_matcher* h = tlsHandle.get();
    if (!h || orbb200_distinctive_descriptors(h, 1, offsets, &flat[0], N, &best, 0, 0) != ORBB200_OK) {
        std::fprintf(stderr, "MapPoint(B200)::ComputeDistinctiveDescriptors: %s\n", orbb200_last_error());
        return;                                   // the previous descriptor stays; no CPU path here
    }
    {
        std::unique_lock<std::mutex> lock(mMutexFeatures);
        mDescriptor = vDescriptors[best].clone();
    }
}

}  // namespace ORB_SLAM2
