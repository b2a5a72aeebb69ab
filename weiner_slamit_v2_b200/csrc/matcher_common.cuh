// matcher_common.cuh -- what the three matcher translation units share: device helpers (descriptor distance, grid
// cell range, the associative best / second-best pair, the sorted top-4 list) and the host-side handle, staging
// allocator and frame-view helpers.  orb_matcher.cu holds the core rows (DescriptorDistance,
// SearchForInitialization, SearchByProjection(F, vpMapPoints, th)), orb_matcher_proj.cu the projection searches that
// compute the projection on the device (LastFrame, KeyFrame relocalisation, Scw, Fuse, SearchBySim3 legs),
// orb_matcher_bow.cu the vocabulary-node searches, the DBoW2 transform and the distinctive-descriptor selection.
#pragma once
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstring>
#include <vector>
#include "common.cuh"

namespace orbb200 {

constexpr int GRID_COLS = 64, GRID_ROWS = 48, GRID_CELLS = GRID_COLS * GRID_ROWS;   // I/Frame.h:40-41
constexpr int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;                          // S/ORBmatcher.cc:37-39

struct FrameDev {           // device-side orbb200_frame_view
    const int* n;
    const float *x, *y;
    const int* octave;
    const float* angle;
    const uint8_t* desc;
    int stride;
};

struct GridGeo { float minX, minY, invW, invH; };

__device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1)
{
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// The same distance with 5 POPC instead of 8: three carry-save adders (2 LOP3 each) compress seven of the eight
// XOR words into one "ones" and three "twos" words.  POPC issues at 16 lanes/clk/SM on sm_100a against 64 for
// LOP3/IADD3 (tools/int_peak.cu, profiles/r1k_int_peak.json), so in a loop that does nothing but distances the
// POPC pipe (8 x 8 cycles per warp) is the bound; this form balances it against the ALU pipe (40 / 40 cycles).
// (inline PTX: left to itself the compiler folds the XORs into the adders and ends up with 20 LOP3 instead of 14)
__device__ __forceinline__ uint32_t lop_xor(uint32_t a, uint32_t b)
{
    uint32_t r; asm("lop3.b32 %0, %1, %2, 0, 0x3c;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t csa_xor3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ uint32_t csa_maj3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ int hamming256_csa(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1)
{
    const uint32_t x0 = lop_xor(a0.x, b0.x), x1 = lop_xor(a0.y, b0.y), x2 = lop_xor(a0.z, b0.z), x3 = lop_xor(a0.w, b0.w);
    const uint32_t x4 = lop_xor(a1.x, b1.x), x5 = lop_xor(a1.y, b1.y), x6 = lop_xor(a1.z, b1.z), x7 = lop_xor(a1.w, b1.w);
    const uint32_t s1 = csa_xor3(x0, x1, x2), c1 = csa_maj3(x0, x1, x2);
    const uint32_t s2 = csa_xor3(x3, x4, x5), c2 = csa_maj3(x3, x4, x5);
    const uint32_t s3 = csa_xor3(s1, s2, x6), c3 = csa_maj3(s1, s2, x6);
    return __popc(s3) + __popc(x7) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
}

// shared-memory accesses by 32-bit shared-space address
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ uint4 lds_v4(uint32_t addr)
{
    uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)); return v;
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// Distance over the first 128 bits only (3 POPC + 6 LOP3): a lower bound of the full distance.
__device__ __forceinline__ int hamming128_lower(const uint4 a0, const uint4 b0)
{
    const uint32_t x0 = lop_xor(a0.x, b0.x), x1 = lop_xor(a0.y, b0.y), x2 = lop_xor(a0.z, b0.z), x3 = lop_xor(a0.w, b0.w);
    return __popc(csa_xor3(x0, x1, x2)) + __popc(x3) + 2 * __popc(csa_maj3(x0, x1, x2));
}

// Cell range of GetFeaturesInArea (S/Frame.cc:452-466); false when the query misses the grid.
__device__ __forceinline__ bool cell_range(const GridGeo& g, float x, float y, float r, int& c0, int& c1, int& r0, int& r1)
{
    c0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, g.minX), r), g.invW)));
    if (c0 >= GRID_COLS) return false;
    c1 = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, g.minX), r), g.invW)));
    if (c1 < 0) return false;
    r0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, g.minY), r), g.invH)));
    if (r0 >= GRID_ROWS) return false;
    r1 = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, g.minY), r), g.invH)));
    if (r1 < 0) return false;
    return true;
}

// Best / second-best exactly as the reference's scan computes them.  The sequential update
//   if (d < best) { second = best; best = d; } else if (d < second) second = d;
// leaves best = smallest and second = second smallest element under the lexicographic key
// (distance, visiting position) -- including which element supplies the "level" payload -- so the
// pair can be reduced associatively across lanes.
struct Top2 { int b, bp, ba, s, sp, sa; };     // best: dist, pos, payload; second: dist, pos, payload
__device__ __forceinline__ void top2_push(Top2& t, int d, int pos, int payload)
{
    if (d < t.b) { t.s = t.b; t.sp = t.bp; t.sa = t.ba; t.b = d; t.bp = pos; t.ba = payload; }
    else if (d < t.s) { t.s = d; t.sp = pos; t.sa = payload; }
}
__device__ __forceinline__ bool key_lt(int d0, int p0, int d1, int p1) { return d0 < d1 || (d0 == d1 && p0 < p1); }
__device__ __forceinline__ Top2 top2_merge(const Top2& a, const Top2& o)
{
    Top2 r;
    if (key_lt(a.b, a.bp, o.b, o.bp)) {
        r.b = a.b; r.bp = a.bp; r.ba = a.ba;
        if (key_lt(a.s, a.sp, o.b, o.bp)) { r.s = a.s; r.sp = a.sp; r.sa = a.sa; } else { r.s = o.b; r.sp = o.bp; r.sa = o.ba; }
    } else {
        r.b = o.b; r.bp = o.bp; r.ba = o.ba;
        if (key_lt(o.s, o.sp, a.b, a.bp)) { r.s = o.s; r.sp = o.sp; r.sa = o.sa; } else { r.s = a.b; r.sp = a.bp; r.sa = a.ba; }
    }
    return r;
}
__device__ __forceinline__ Top2 top2_warp_reduce(Top2 t)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        Top2 o;
        o.b = __shfl_xor_sync(0xffffffffu, t.b, d); o.bp = __shfl_xor_sync(0xffffffffu, t.bp, d);
        o.ba = __shfl_xor_sync(0xffffffffu, t.ba, d); o.s = __shfl_xor_sync(0xffffffffu, t.s, d);
        o.sp = __shfl_xor_sync(0xffffffffu, t.sp, d); o.sa = __shfl_xor_sync(0xffffffffu, t.sa, d);
        t = top2_merge(t, o);
    }
    return t;
}

// Sorted insertion of a key into a 4-entry ascending list (keys are unique: they embed the position).
__device__ __forceinline__ void top4_insert(uint4& t, uint32_t k)
{
    uint32_t m;
    m = min(t.x, k); k = max(t.x, k); t.x = m;
    m = min(t.y, k); k = max(t.y, k); t.y = m;
    m = min(t.z, k); k = max(t.z, k); t.z = m;
    t.w = min(t.w, k);
}

// The speculative greedy resolve (k_search_init / k_search_proj / k_search_last): each lane has decided its own query
// against the current state; `w` is the keypoint whose state the lane would change (-1: none) and kid[] the <= 4
// keypoints whose state it read.  A lane is final when no EARLIER lane writes one of the keypoints it read.  Returns
// the first lane that is not final (32 if all are): writers publish their lane number in `claim` (one byte per
// keypoint, 0xff when idle; the lowest lane of several writers of one keypoint), everybody looks its reads up, and
// the claims are withdrawn again -- a constant number of steps instead of one broadcast per writer.
__device__ __forceinline__ int clean_prefix(uint8_t* claim, int lane, int w, const int (&kid)[4], bool blockedSelf)
{
    const unsigned same = __match_any_sync(0xffffffffu, w >= 0 ? w : -1 - lane);
    const bool leader = w >= 0 && lane == __ffs(same) - 1;
    if (leader) claim[w] = (uint8_t)lane;
    __syncwarp();
    bool blocked = blockedSelf;
#pragma unroll
    for (int e = 0; e < 4; e++)
        if (kid[e] >= 0 && (int)claim[kid[e]] < lane) blocked = true;
    __syncwarp();
    if (leader) claim[w] = 0xff;
    const unsigned bm = __ballot_sync(0xffffffffu, blocked);
    return bm ? __ffs(bm) - 1 : 32;
}

// Frame::AssignFeaturesToGrid as CSR for `items` frames on stream st (kernel in orb_matcher.cu)
int launch_build_grid(const FrameDev& f, const GridGeo& g, int* cellStart, int* cellItems, uint4* cellRec, int items, cudaStream_t st);

}  // namespace orbb200

// =========================================================================================
// host side
// =========================================================================================
using orbb200::FrameDev; using orbb200::GridGeo; using orbb200::GRID_COLS; using orbb200::GRID_ROWS; using orbb200::GRID_CELLS;
using orbb200::set_error; using orbb200::align_up;
struct orbb200_matcher {
    int maxItems, maxPoints, device, lastLaunches;
    cudaStream_t stream;
    int *cellStart, *cellItems, *scratchA, *scratchB, *scratchC, *topkCount;
    uint4 *topk, *topkIdx;
    uint4* cellRec;         // items x maxPoints x 3: the grid's keypoints in CSR order, {x, y, index, octave} + descriptor (k_build_grid)
    std::vector<void*> allocs;
    // staging for host-pointer calls
    uint8_t* stage; size_t stageBytes;
};

// bump allocator over one device staging block for host-pointer calls
struct Stager {
    orbb200_matcher* m; size_t off; cudaStream_t st;
    int reserve(size_t bytes)
    {
        if (bytes > m->stageBytes) {
            if (m->stage) cudaFree(m->stage);
            m->stage = nullptr; m->stageBytes = 0;
            ORB_CUDA(cudaMalloc((void**)&m->stage, bytes));
            m->stageBytes = bytes;
        }
        off = 0;
        return ORBB200_OK;
    }
    template <typename T> int up(const T* host, size_t count, const T** dev)
    {
        if (!host) { *dev = nullptr; return ORBB200_OK; }
        T* d = reinterpret_cast<T*>(m->stage + off);
        off += align_up(count * sizeof(T), 256);
        ORB_CUDA(cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, st));
        *dev = d;
        return ORBB200_OK;
    }
    template <typename T> T* out(size_t count)
    {
        T* d = reinterpret_cast<T*>(m->stage + off);
        off += align_up(count * sizeof(T), 256);
        return d;
    }
};
static inline size_t pad(size_t b) { return align_up(b, 256); }


static inline GridGeo grid_geo(const float* bounds)
{
    GridGeo g;
    g.minX = bounds[0]; g.minY = bounds[1];                           // Frame::mnMinX/Y (S/Frame.cc:561-589)
    g.invW = (float)GRID_COLS / (bounds[2] - bounds[0]);               // mfGridElementWidthInv (S/Frame.cc:317-318)
    g.invH = (float)GRID_ROWS / (bounds[3] - bounds[1]);
    return g;
}

static inline int check_view(const orbb200_matcher* m, int items, int stride, const char* what)
{
    if (items < 1 || items > m->maxItems) { set_error("%s: items %d outside 1..%d", what, items, m->maxItems); return ORBB200_EINVAL; }
    if (stride < 1 || stride > m->maxPoints) { set_error("%s: stride %d outside 1..%d", what, stride, m->maxPoints); return ORBB200_EINVAL; }
    if (stride >= (1 << 18)) { set_error("%s: more than 262143 points per item", what); return ORBB200_EINVAL; }
    return ORBB200_OK;
}

static inline int upload_frame(Stager& s, const orbb200_frame_view* v, int items, FrameDev* d, bool needAngle)
{
    const size_t np = (size_t)items * v->stride;
    int rc;
    d->stride = v->stride;
    if ((rc = s.up(v->n, items, &d->n))) return rc;
    if ((rc = s.up(v->x, np, &d->x))) return rc;
    if ((rc = s.up(v->y, np, &d->y))) return rc;
    if ((rc = s.up(v->octave, np, &d->octave))) return rc;
    if ((rc = s.up(needAngle ? v->angle : nullptr, np, &d->angle))) return rc;
    if ((rc = s.up(v->desc, np * 32, &d->desc))) return rc;
    return ORBB200_OK;
}
static inline size_t frame_bytes(const orbb200_frame_view* v, int items)
{
    const size_t np = (size_t)items * v->stride;
    return pad(items * 4) + 4 * pad(np * 4) + pad(np * 32);
}
static inline FrameDev as_dev(const orbb200_frame_view* v)
{
    FrameDev d;
    d.n = v->n; d.x = v->x; d.y = v->y; d.octave = v->octave; d.angle = v->angle; d.desc = v->desc; d.stride = v->stride;
    return d;
}
