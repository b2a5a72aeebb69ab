// ref_harness.cc -- C entry points around the reference's own ORBextractor (TEST
// INFRASTRUCTURE).  ORB_SLAM2/src/ORBextractor.cc is compiled unmodified, straight from
// /root/reference, against mini-cv; this file only (1) subclasses it to reach the protected
// stage methods, (2) installs a monotonic bump allocator so the quadtree's pointer-valued
// tie-break (ORBextractor.cc:694-698) is reproducible: address order == creation order.
#include <sys/mman.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>
#include "ORBextractor.h"
#include "orb_oracle.h"

// ---- monotonic per-thread arena behind operator new -------------------------------------
// (-DREF_DEFAULT_MALLOC builds the same library WITHOUT it: libref_orb_malloc.so, the reference under glibc's allocator,
// for the tie-break sensitivity report tools/tiebreak_report.py -> profiles/r2_tiebreak.json)
#ifndef REF_DEFAULT_MALLOC
namespace {
struct Arena {
    char* base; size_t top, cap;
    int direction;   // 0 ascending addresses (canonical), 1 descending
};
thread_local Arena g_arena = {nullptr, 0, 0, 0};
int g_direction = 0;
const size_t ARENA_BYTES = (size_t)8 << 30;   // virtual, committed lazily

void* arena_alloc(size_t n)
{
    Arena& a = g_arena;
    if (!a.base) {
        void* p = mmap(nullptr, ARENA_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { std::fprintf(stderr, "ref_harness: arena mmap failed\n"); std::abort(); }
        a.base = (char*)p; a.top = 0; a.cap = ARENA_BYTES;
    }
    n = (n + 15) & ~(size_t)15;
    if (n == 0) n = 16;
    if (a.top + n > a.cap) { std::fprintf(stderr, "ref_harness: arena exhausted\n"); std::abort(); }
    a.top += n;
    return g_direction == 0 ? a.base + (a.top - n) : a.base + (a.cap - a.top);
}
}  // namespace

void* operator new(size_t n) { return arena_alloc(n); }
void* operator new[](size_t n) { return arena_alloc(n); }
void operator delete(void*) noexcept {}
void operator delete[](void*) noexcept {}
void operator delete(void*, size_t) noexcept {}
void operator delete[](void*, size_t) noexcept {}
#else
namespace { int g_direction = 0; struct NoArena { size_t top; }; thread_local NoArena g_arena = {0}; }   // marks are no-ops
#endif

namespace {
class RefExtractor : public ORB_SLAM2::ORBextractor {
public:
    RefExtractor(int n, float s, int l, int a, int b) : ORB_SLAM2::ORBextractor(n, s, l, a, b) {}
    std::vector<cv::KeyPoint> octree(const std::vector<cv::KeyPoint>& v, int minX, int maxX, int minY, int maxY, int N, int level)
    { return DistributeOctTree(v, minX, maxX, minY, maxY, N, level); }
    const std::vector<int>& perLevel() const { return mnFeaturesPerLevel; }
    const std::vector<int>& uMax() const { return umax; }
    const std::vector<cv::Point>& pat() const { return pattern; }
};

void to_c(const cv::KeyPoint& k, orc_keypoint* o)
{
    o->x = k.pt.x; o->y = k.pt.y; o->size = k.size; o->angle = k.angle; o->response = k.response;
    o->octave = k.octave; o->class_id = k.class_id;
}
}  // namespace

extern "C" {

void ref_set_alloc_direction(int d) { g_direction = d; }

void* ref_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh)
{
    return new RefExtractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void ref_extractor_destroy(void* h) { delete (RefExtractor*)h; }

void ref_tables(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* per_level, int* umax16, int* pattern1024)
{
    RefExtractor* e = (RefExtractor*)h;
    int L = e->GetLevels();
    std::vector<float> a = e->GetScaleFactors(), b = e->GetInverseScaleFactors(), c = e->GetScaleSigmaSquares(), d = e->GetInverseScaleSigmaSquares();
    for (int i = 0; i < L; i++) { scale[i] = a[i]; inv_scale[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i]; per_level[i] = e->perLevel()[i]; }
    for (int i = 0; i < 16; i++) umax16[i] = e->uMax()[i];
    for (int i = 0; i < 512; i++) { pattern1024[2 * i] = e->pat()[i].x; pattern1024[2 * i + 1] = e->pat()[i].y; }
}

// ORBextractor::operator() on one frame.  Returns the keypoint count, or -1 if cap is too small.
int ref_extract(void* h, const uint8_t* img, int w, int h_, int stride, orc_keypoint* kps, uint8_t* desc, int cap)
{
    RefExtractor* e = (RefExtractor*)h;
    size_t mark = g_arena.top;
    int n;
    {
        cv::Mat image(h_, w, CV_8UC1, (void*)img, (size_t)stride);
        std::vector<cv::KeyPoint> keys;
        cv::Mat descriptors;
        (*e)(image, cv::Mat(), keys, descriptors);
        n = (int)keys.size();
        if (n <= cap) {
            for (int i = 0; i < n; i++) to_c(keys[i], &kps[i]);
            for (int i = 0; i < n; i++) std::memcpy(desc + 32 * (size_t)i, descriptors.ptr(i), 32);
        } else {
            n = -1;
        }
    }
    g_arena.top = mark;   // everything the call allocated is dead now
    return n;
}

// pyramid level of the most recent call, border (19 px) included when with_border != 0
int ref_level_size(void* h, int level, int* w, int* hh)
{
    RefExtractor* e = (RefExtractor*)h;
    *w = e->mvImagePyramid[level].cols; *hh = e->mvImagePyramid[level].rows;
    return 0;
}
void ref_level_pixels(void* h, int level, int with_border, uint8_t* dst)
{
    RefExtractor* e = (RefExtractor*)h;
    const cv::Mat& m = e->mvImagePyramid[level];
    int b = with_border ? 19 : 0;
    for (int y = -b; y < m.rows + b; y++)
        std::memcpy(dst + (size_t)(y + b) * (m.cols + 2 * b), m.data + (ptrdiff_t)y * (ptrdiff_t)m.step - b, m.cols + 2 * b);
}

// ORBextractor::DistributeOctTree alone
int ref_distribute_octree(void* h, const orc_keypoint* in, int n, int minX, int maxX, int minY, int maxY, int N, int level,
                          orc_keypoint* out, int cap)
{
    RefExtractor* e = (RefExtractor*)h;
    size_t mark = g_arena.top;
    int nout;
    {
        std::vector<cv::KeyPoint> v;
        v.reserve(n);
        for (int i = 0; i < n; i++) v.push_back(cv::KeyPoint(in[i].x, in[i].y, in[i].size, in[i].angle, in[i].response, in[i].octave, in[i].class_id));
        std::vector<cv::KeyPoint> r = e->octree(v, minX, maxX, minY, maxY, N, level);
        nout = (int)r.size();
        if (nout <= cap) for (int i = 0; i < nout; i++) to_c(r[i], &out[i]); else nout = -1;
    }
    g_arena.top = mark;
    return nout;
}

// CPU baseline: T threads, one extractor each (instances are not re-entrant), disjoint
// contiguous frame ranges.  counts[i] receives the keypoint count of frame i.
int ref_extract_batch_mt(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh,
                         const uint8_t* frames, int nframes, int w, int h_, int threads, int* counts)
{
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) {
        pool.emplace_back([=]() {
            RefExtractor ex(nfeatures, scaleFactor, nlevels, iniTh, minTh);
            int lo = (int)((long)nframes * t / threads), hi = (int)((long)nframes * (t + 1) / threads);
            for (int f = lo; f < hi; f++) {
                size_t mark = g_arena.top;
                {
                    cv::Mat image(h_, w, CV_8UC1, (void*)(frames + (size_t)f * w * h_), (size_t)w);
                    std::vector<cv::KeyPoint> keys;
                    cv::Mat descriptors;
                    ex(image, cv::Mat(), keys, descriptors);
                    counts[f] = (int)keys.size();
                }
                g_arena.top = mark;
            }
        });
    }
    for (auto& th : pool) th.join();
    return 0;
}
}
