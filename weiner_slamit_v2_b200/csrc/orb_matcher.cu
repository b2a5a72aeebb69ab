// orb_matcher.cu -- ORBmatcher's Hamming search for batches of frame pairs / frames on B200.
//
// Reference path (S/ = oRB_SLAM2_Android/src/main/jni/ORB_SLAM2/src/):
//   DescriptorDistance S/ORBmatcher.cc:1651-1667, SearchForInitialization :409-524,
//   SearchByProjection(Frame&, vector<MapPoint*>&, th) :47-131, ComputeThreeMaxima :1605-1646,
//   Frame::AssignFeaturesToGrid / GetFeaturesInArea / PosInGrid S/Frame.cc:336-357, 447-517.
//
// Both searches are greedy and order dependent inside one frame (pair): an accepted match changes
// what later queries may take (vMatchedDistance :448, mvpMapPoints :89-91,:125).  Items (pairs /
// frames) are independent, so the device mapping is: items in parallel, one warp per item walking
// its queries in the reference's order, the 32 lanes sharing each query's candidate scan
// (XOR + __popc on 256-bit descriptors) and a warp-shuffle best / second-best reduction whose
// tie-break reproduces "first candidate in GetFeaturesInArea order wins" (candidates are visited
// in CSR order = ix, iy, insertion order; ties go to the lower CSR position).
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstring>
#include <vector>
#include <math_constants.h>
#include "common.cuh"
#include "../../include/orb_b200_logf.inc"

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

// ---- DescriptorDistance for n independent pairs ------------------------------------------
__global__ void k_distance(const uint4* a, const uint4* b, int n, int* dist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dist[i] = hamming256(__ldg(a + 2 * i), __ldg(a + 2 * i + 1), __ldg(b + 2 * i), __ldg(b + 2 * i + 1));
}

// ---- Frame::AssignFeaturesToGrid as CSR (one CTA per frame) --------------------------------
// cell = ix*48+iy with ix,iy = roundf((pt - min) * inv) (PosInGrid, S/Frame.cc:505-517); points whose
// cell falls outside the 64x48 grid are not indexed.  Inside a cell, indices ascend (push_back order).
__global__ void __launch_bounds__(256) k_build_grid(FrameDev f, GridGeo g, int* cellStart, int* cellItems)
{
    __shared__ int cnt[GRID_CELLS + 1];
    __shared__ int wsum[9];
    const int item = blockIdx.x, tid = threadIdx.x;
    const int n = min(f.n[item], f.stride);
    const float* kx = f.x + (size_t)item * f.stride;
    const float* ky = f.y + (size_t)item * f.stride;
    int* cs = cellStart + (size_t)item * (GRID_CELLS + 1);
    int* ci = cellItems + (size_t)item * f.stride;
    for (int c = tid; c <= GRID_CELLS; c += 256) cnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const int px = (int)roundf(__fmul_rn(__fsub_rn(kx[i], g.minX), g.invW));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(ky[i], g.minY), g.invH));
        if (px >= 0 && px < GRID_COLS && py >= 0 && py < GRID_ROWS) atomicAdd(&cnt[px * GRID_ROWS + py], 1);
    }
    __syncthreads();
    // exclusive scan of 3072 counts: 12 per thread
    int local[12], sum = 0;
#pragma unroll
    for (int j = 0; j < 12; j++) { local[j] = cnt[tid * 12 + j]; sum += local[j]; }
    int x = sum;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (tid == 0) { int a = 0; for (int w = 0; w < 8; w++) { const int t = wsum[w]; wsum[w] = a; a += t; } wsum[8] = a; }
    __syncthreads();
    int run = x - sum + wsum[warp];
#pragma unroll
    for (int j = 0; j < 12; j++) { cnt[tid * 12 + j] = run; cs[tid * 12 + j] = run; run += local[j]; }
    if (tid == 255) { cs[GRID_CELLS] = run; }
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const int px = (int)roundf(__fmul_rn(__fsub_rn(kx[i], g.minX), g.invW));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(ky[i], g.minY), g.invH));
        if (px >= 0 && px < GRID_COLS && py >= 0 && py < GRID_ROWS) ci[atomicAdd(&cnt[px * GRID_ROWS + py], 1)] = i;
    }
    __syncthreads();
    // restore insertion (ascending index) order inside every cell
    for (int c = tid; c < GRID_CELLS; c += 256) {
        const int s = cs[c], e = cnt[c];
        for (int a = s + 1; a < e; a++) {
            const int v = ci[a];
            int b = a - 1;
            while (b >= s && ci[b] > v) { ci[b + 1] = ci[b]; b--; }
            ci[b + 1] = v;
        }
    }
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

// ---- SearchForInitialization: one warp per frame pair (S/ORBmatcher.cc:409-524) -------------
struct InitParams {
    FrameDev f1, f2;
    GridGeo g;
    const int* cellStart;     // items x (GRID_CELLS+1), F2 grid
    const int* cellItems;     // items x f2.stride
    int* matches21;           // items x f2.stride scratch (vnMatches21)
    int* histBin;             // items x f1.stride scratch: rotation bin of i1 when it was accepted, else -1
    float* prevMatched;       // items x f1.stride x 2, in/out
    int* matches12;           // items x f1.stride
    int* nmatches;            // items
    uint4* topk;              // items x f1.stride: the 4 best candidates of every query, key = dist << 20 | CSR position
    int* topkCount;           // items x f1.stride: number of candidates that passed the static tests
    uint4* topkIdx;           // items x f1.stride: F2 keypoint index of each of the 4 entries
    int items, window, checkOri;
    float nnratio;
};

// Sorted insertion of a key into a 4-entry ascending list (keys are unique: they embed the position).
__device__ __forceinline__ void top4_insert(uint4& t, uint32_t k)
{
    uint32_t m;
    m = min(t.x, k); k = max(t.x, k); t.x = m;
    m = min(t.y, k); k = max(t.y, k); t.y = m;
    m = min(t.z, k); k = max(t.z, k); t.z = m;
    t.w = min(t.w, k);
}

// ---- SearchForInitialization, phase A: everything that does not depend on the greedy state ----
// One THREAD per F1 keypoint (query); the F2 keypoints are streamed through shared memory in CSR order
// (= GetFeaturesInArea visiting order) and every lane tests the same candidate at the same time, so the
// candidate's descriptor is one broadcast load.  A candidate that passes the static tests of
// GetFeaturesInArea (cell range, level, window) costs 8 XOR + 8 POPC; the 4 smallest keys
// (distance, visiting position) and the candidate count are kept per query.
constexpr int TOPK_CHUNK = 1024;
struct CandMeta { float x, y; int meta; };     // meta = octave | cx << 16 | cy << 24

__global__ void __launch_bounds__(128) k_init_topk(const InitParams P)
{
    __shared__ __align__(16) uint4 s_desc[TOPK_CHUNK * 2];
    __shared__ CandMeta s_meta[TOPK_CHUNK];
    const int item = blockIdx.y, tid = threadIdx.x;
    const int n1 = min(P.f1.n[item], P.f1.stride), n2 = min(P.f2.n[item], P.f2.stride);
    if ((int)blockIdx.x * 128 >= n1) return;
    const int q = blockIdx.x * 128 + tid;
    const float* k2x = P.f2.x + (size_t)item * P.f2.stride;
    const float* k2y = P.f2.y + (size_t)item * P.f2.stride;
    const int* oct2 = P.f2.octave + (size_t)item * P.f2.stride;
    const uint4* d2 = reinterpret_cast<const uint4*>(P.f2.desc + (size_t)item * P.f2.stride * 32);
    const int* cs = P.cellStart + (size_t)item * (GRID_CELLS + 1);
    const int* ci = P.cellItems + (size_t)item * P.f2.stride;
    const int ngrid = cs[GRID_CELLS];                          // keypoints that are in the grid at all

    bool active = q < n1;
    int level1 = 0, c0 = 0, c1 = -1, r0 = 0, r1 = -1;
    float qx = 0.f, qy = 0.f;
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
    const float r = (float)P.window;
    if (active) {
        level1 = P.f1.octave[(size_t)item * P.f1.stride + q];
        const float* prev = P.prevMatched + ((size_t)item * P.f1.stride + q) * 2;
        qx = prev[0]; qy = prev[1];
        active = level1 <= 0 && cell_range(P.g, qx, qy, r, c0, c1, r0, r1);
        const uint4* d1 = reinterpret_cast<const uint4*>(P.f1.desc + ((size_t)item * P.f1.stride + q) * 32);
        a0 = __ldg(d1); a1 = __ldg(d1 + 1);
    }
    uint4 best = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    int count = 0;
    for (int base = 0; base < ngrid; base += TOPK_CHUNK) {
        const int nc = min(TOPK_CHUNK, ngrid - base);
        __syncthreads();
        for (int c = tid; c < nc; c += 128) {
            const int i2 = ci[base + c];
            const float x = k2x[i2], y = k2y[i2];
            const int cx = (int)roundf(__fmul_rn(__fsub_rn(x, P.g.minX), P.g.invW));
            const int cy = (int)roundf(__fmul_rn(__fsub_rn(y, P.g.minY), P.g.invH));
            s_meta[c].x = x; s_meta[c].y = y;
            s_meta[c].meta = (oct2[i2] & 0xffff) | (cx << 16) | (cy << 24);
            s_desc[2 * c] = __ldg(d2 + 2 * i2);
            s_desc[2 * c + 1] = __ldg(d2 + 2 * i2 + 1);
        }
        __syncthreads();
        if (active) {
            for (int c = 0; c < nc; c++) {
                const CandMeta m = s_meta[c];
                const int o2 = (short)(m.meta & 0xffff), cx = (m.meta >> 16) & 0xff, cy = (m.meta >> 24) & 0xff;
                if (cx < c0 || cx > c1 || cy < r0 || cy > r1) continue;
                if (o2 < level1 || (level1 >= 0 && o2 > level1)) continue;             // Frame.cc:468-485
                if (!(fabsf(__fsub_rn(m.x, qx)) < r && fabsf(__fsub_rn(m.y, qy)) < r)) continue;
                const int dist = hamming256(a0, a1, s_desc[2 * c], s_desc[2 * c + 1]);
                top4_insert(best, ((uint32_t)dist << 20) | (uint32_t)(base + c));
                count++;
            }
        }
    }
    if (q < n1) {
        P.topk[(size_t)item * P.f1.stride + q] = best;
        P.topkCount[(size_t)item * P.f1.stride + q] = active ? count : -1;
        if (active && count > 0) {
            const uint32_t k[4] = {best.x, best.y, best.z, best.w};
            uint32_t id[4];
#pragma unroll
            for (int j = 0; j < 4; j++) id[j] = j < count ? (uint32_t)ci[k[j] & 0xfffffu] : 0u;
            P.topkIdx[(size_t)item * P.f1.stride + q] = make_uint4(id[0], id[1], id[2], id[3]);
        }
    }
}

// ---- SearchForInitialization, phase B: the greedy pass, one warp per frame pair ----------------
__global__ void __launch_bounds__(128) k_search_init(const InitParams P)
{
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= P.items) return;
    const int n1 = min(P.f1.n[item], P.f1.stride), n2 = min(P.f2.n[item], P.f2.stride);
    const int* oct1 = P.f1.octave + (size_t)item * P.f1.stride;
    const float* ang1 = P.f1.angle + (size_t)item * P.f1.stride;
    const uint4* d1 = reinterpret_cast<const uint4*>(P.f1.desc + (size_t)item * P.f1.stride * 32);
    const float* k2x = P.f2.x + (size_t)item * P.f2.stride;
    const float* k2y = P.f2.y + (size_t)item * P.f2.stride;
    const int* oct2 = P.f2.octave + (size_t)item * P.f2.stride;
    const float* ang2 = P.f2.angle + (size_t)item * P.f2.stride;
    const uint4* d2 = reinterpret_cast<const uint4*>(P.f2.desc + (size_t)item * P.f2.stride * 32);
    const int* cs = P.cellStart + (size_t)item * (GRID_CELLS + 1);
    const int* ci = P.cellItems + (size_t)item * P.f2.stride;
    int* m21 = P.matches21 + (size_t)item * P.f2.stride;
    int* hbin = P.histBin + (size_t)item * P.f1.stride;
    float* prev = P.prevMatched + (size_t)item * P.f1.stride * 2;
    int* m12 = P.matches12 + (size_t)item * P.f1.stride;

    // vMatchedDistance (:419) lives in shared memory as 16 bits (0xffff = INT_MAX; real values are <= TH_LOW)
    extern __shared__ uint16_t s_vmd_all[];
    uint16_t* vmd = s_vmd_all + (size_t)(threadIdx.x >> 5) * ((P.f2.stride + 7) & ~7);
    for (int i = lane; i < n2; i += 32) { vmd[i] = 0xffff; m21[i] = -1; }
    for (int i = lane; i < n1; i += 32) { m12[i] = -1; hbin[i] = -1; }
    __syncwarp();

    const float r = (float)P.window;
    const uint4* topk = P.topk + (size_t)item * P.f1.stride;
    const uint4* topkIdx = P.topkIdx + (size_t)item * P.f1.stride;
    const int* topkCount = P.topkCount + (size_t)item * P.f1.stride;
    // lane j loads the list of query base + j (coalesced, one latency per 32 queries); the warp then walks the
    // queries in order, broadcasting each list by shuffle
    for (int base = 0; base < n1; base += 32) {
      const int mine = base + lane;
      int cntL = -1;
      uint4 kkL = make_uint4(0, 0, 0, 0), idL = kkL;
      if (mine < n1) { cntL = topkCount[mine]; kkL = topk[mine]; idL = topkIdx[mine]; }
      const int jEnd = min(32, n1 - base);
      for (int jq = 0; jq < jEnd; jq++) {
        const int i1 = base + jq;
        const int cnt = __shfl_sync(0xffffffffu, cntL, jq);
        if (cnt <= 0) continue;            // octave > 0 (:425-427), query outside the grid, or no candidate (:431)
        const uint4 kk = make_uint4(__shfl_sync(0xffffffffu, kkL.x, jq), __shfl_sync(0xffffffffu, kkL.y, jq),
                                    __shfl_sync(0xffffffffu, kkL.z, jq), __shfl_sync(0xffffffffu, kkL.w, jq));
        const uint4 idv = make_uint4(__shfl_sync(0xffffffffu, idL.x, jq), __shfl_sync(0xffffffffu, idL.y, jq),
                                     __shfl_sync(0xffffffffu, idL.z, jq), __shfl_sync(0xffffffffu, idL.w, jq));
        const int level1 = oct1[i1];
        Top2 t = {INT_MAX, INT_MAX, -1, INT_MAX, INT_MAX, -1};
        // Fast path (every lane redundantly): walk the query's 4 best static candidates in visiting order and
        // drop those a previous query already holds at a distance <= ours (:448).  The first two survivors are
        // best / second-best.  That is conclusive when two survive, when the list holds every candidate, or
        // when the best survivor already fails TH_LOW; otherwise fall back to the full scan below.
        bool resolved;
        {
            const uint32_t key[4] = {kk.x, kk.y, kk.z, kk.w};
            const uint32_t kid[4] = {idv.x, idv.y, idv.z, idv.w};
            int found = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (found < 2 && j < cnt) {
                    const int dist = (int)(key[j] >> 20), pos = (int)(key[j] & 0xfffffu);
                    if (!((int)vmd[kid[j]] <= dist)) {
                        if (found == 0) { t.b = dist; t.bp = pos; } else { t.s = dist; t.sp = pos; }
                        found++;
                    }
                }
            }
            // every candidate beyond the list is at least as far as its last entry (d3)
            const int d3 = (int)(key[3] >> 20);
            resolved = found == 2 || cnt <= 4 || (found == 1 && t.b > TH_LOW) || (found == 0 && d3 > TH_LOW);
            if (!resolved && found == 1 && (float)t.b < __fmul_rn((float)d3, P.nnratio)) {
                t.s = d3;                  // the true second-best is >= d3: the ratio test passes either way
                resolved = true;
            }
        }
        if (!resolved) {
            t = Top2{INT_MAX, INT_MAX, -1, INT_MAX, INT_MAX, -1};
            int c0, c1, r0, r1;
            cell_range(P.g, prev[2 * i1], prev[2 * i1 + 1], r, c0, c1, r0, r1);
            const float qx = prev[2 * i1], qy = prev[2 * i1 + 1];
            const uint4 a0 = __ldg(d1 + 2 * i1), a1 = __ldg(d1 + 2 * i1 + 1);
            const bool fullRows = (r0 == 0 && r1 == GRID_ROWS - 1);
            const int ncol = fullRows ? 1 : (c1 - c0 + 1);
            for (int c = 0; c < ncol; c++) {
                const int s = fullRows ? cs[c0 * GRID_ROWS] : cs[(c0 + c) * GRID_ROWS + r0];
                const int e = fullRows ? cs[(c1 + 1) * GRID_ROWS] : cs[(c0 + c) * GRID_ROWS + r1 + 1];
                for (int p = s + lane; p < e; p += 32) {
                    const int i2 = ci[p];
                    // level filter: minLevel = maxLevel = level1 (:429, Frame.cc:468-485)
                    const int o2 = oct2[i2];
                    if (o2 < level1 || (level1 >= 0 && o2 > level1)) continue;
                    if (!(fabsf(__fsub_rn(k2x[i2], qx)) < r && fabsf(__fsub_rn(k2y[i2], qy)) < r)) continue;
                    const int dist = hamming256(a0, a1, __ldg(d2 + 2 * i2), __ldg(d2 + 2 * i2 + 1));
                    if ((int)vmd[i2] <= dist) continue;                                // :448
                    top2_push(t, dist, p, 0);
                }
            }
            t = top2_warp_reduce(t);
        }
        if (t.b <= TH_LOW && (float)t.b < __fmul_rn((float)t.s, P.nnratio)) {     // :463-465
            if (lane == 0) {
                const int best2 = ci[t.bp];
                const int old = m21[best2];
                if (old >= 0) m12[old] = -1;                                         // :467-471
                m12[i1] = best2;
                m21[best2] = i1;
                vmd[best2] = (uint16_t)t.b;
                if (P.checkOri) {
                    float rot = __fsub_rn(ang1[i1], ang2[best2]);
                    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                    int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));       // sic (:417,:482)
                    if (bin == HISTO_LENGTH) bin = 0;
                    hbin[i1] = bin;
                }
            }
        }
        __syncwarp();
      }
    }
    // The running count (nmatches++ / nmatches-- at :470,:475) equals the number of live entries of
    // vnMatches12 at this point; recount instead of tracking the decrements.
    __syncwarp();
    int live = 0;
    for (int i = lane; i < n1; i += 32) live += m12[i] >= 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) live += __shfl_xor_sync(0xffffffffu, live, d);
    int nmatches = live;

    if (P.checkOri) {
        // histogram sizes count every acceptance, including matches stolen later (:479-486)
        int sizes = 0;     // lane b holds the size of bin b (30 bins)
        for (int i = 0; i < n1; i += 32) {
            const int b = (i + lane < n1) ? hbin[i + lane] : -1;
            for (int q = 0; q < HISTO_LENGTH; q++) {
                const unsigned m = __ballot_sync(0xffffffffu, b == q);
                if (lane == q) sizes += __popc(m);
            }
        }
        // ComputeThreeMaxima (:1605-1646), evaluated redundantly by every lane
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int q = 0; q < HISTO_LENGTH; q++) {
            const int s = __shfl_sync(0xffffffffu, sizes, q);
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = q; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = q; }
            else if (s > max3) { max3 = s; ind3 = q; }
        }
        if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
        else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
        int removed = 0;
        for (int i = lane; i < n1; i += 32) {
            const int b = hbin[i];
            if (b >= 0 && b != ind1 && b != ind2 && b != ind3 && m12[i] >= 0) { m12[i] = -1; removed++; }   // :493-516
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, d);
        nmatches -= removed;
    }
    __syncwarp();
    for (int i = lane; i < n1; i += 32)                                                    // :519-521
        if (m12[i] >= 0) { prev[2 * i] = k2x[m12[i]]; prev[2 * i + 1] = k2y[m12[i]]; }
    if (lane == 0) P.nmatches[item] = nmatches;
}

// ---- SearchByProjection: one warp per frame (S/ORBmatcher.cc:47-131) ------------------------
struct ProjParams {
    FrameDev f;
    const float* uRight;       // items x f.stride or NULL
    GridGeo g;
    const int* cellStart;
    const int* cellItems;
    const int* mpN;
    const uint8_t *mpInView, *mpBad;
    const float *mpX, *mpY, *mpXR;
    const int* mpLevel;
    const float* mpViewCos;
    const uint8_t* mpDesc;
    const int* mpObs;
    int mpStride;
    int* kpMp;                 // items x f.stride, in/out
    const int* kpMpObs;        // items x f.stride or NULL
    const float* scaleFactors;
    int nlevels;
    int* nmatches;
    uint4* topk;               // items x mpStride: 4 best static candidates, key = dist << 23 | CSR position << 5 | octave
    int* topkCount;            // items x mpStride
    uint4* topkIdx;            // items x mpStride: keypoint index of each of the 4 entries (saves the cellItems hop in phase B)
    int items;
    float nnratio, th;
};

// ---- SearchByProjection, phase A: one THREAD per map point, everything that does not depend on the
// assignments made during the call: grid window, level and stereo tests, keypoints that hold a map
// point with observations from the start (those are never re-assigned), Hamming distance; the 4
// smallest keys (distance, visiting position) and the candidate count are kept per map point.
__global__ void __launch_bounds__(128) k_proj_topk(const ProjParams P)
{
    const int item = blockIdx.y;
    const int nmp = min(P.mpN[item], P.mpStride);
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= nmp) return;
    const size_t mo = (size_t)item * P.mpStride;
    uint4 best = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    const int* ciw = P.cellItems + (size_t)item * P.f.stride;
    int count = -1;
    if (P.mpInView[mo + i] && !P.mpBad[mo + i]) {                                    // :56-60
        const float* kx = P.f.x + (size_t)item * P.f.stride;
        const float* ky = P.f.y + (size_t)item * P.f.stride;
        const int* koct = P.f.octave + (size_t)item * P.f.stride;
        const uint4* kd = reinterpret_cast<const uint4*>(P.f.desc + (size_t)item * P.f.stride * 32);
        const float* ur = P.uRight ? P.uRight + (size_t)item * P.f.stride : nullptr;
        const int* cs = P.cellStart + (size_t)item * (GRID_CELLS + 1);
        const int* ci = P.cellItems + (size_t)item * P.f.stride;
        const int* kpmp = P.kpMp + (size_t)item * P.f.stride;
        const int* kpobs = P.kpMpObs ? P.kpMpObs + (size_t)item * P.f.stride : nullptr;
        const int lvl = P.mpLevel[mo + i];
        float r = ((double)P.mpViewCos[mo + i] > 0.998) ? 2.5f : 4.0f;               // :133-139
        if (P.th != 1.0f) r = __fmul_rn(r, P.th);
        const float rs = __fmul_rn(r, P.scaleFactors[lvl]);
        const float qx = P.mpX[mo + i], qy = P.mpY[mo + i], qxr = P.mpXR[mo + i];
        int c0, c1, r0, r1;
        if (cell_range(P.g, qx, qy, rs, c0, c1, r0, r1)) {
            count = 0;
            const int minLevel = lvl - 1, maxLevel = lvl;
            const bool check = (minLevel > 0) || (maxLevel >= 0);
            const uint4* md = reinterpret_cast<const uint4*>(P.mpDesc + (mo + i) * 32);
            const uint4 a0 = __ldg(md), a1 = __ldg(md + 1);
            for (int c = c0; c <= c1; c++) {
                const int s = cs[c * GRID_ROWS + r0], e = cs[c * GRID_ROWS + r1 + 1];
                for (int p = s; p < e; p++) {
                    const int idx = ci[p];
                    const int o = koct[idx];
                    if (check) {
                        if (o < minLevel) continue;
                        if (maxLevel >= 0 && o > maxLevel) continue;
                    }
                    if (!(fabsf(__fsub_rn(kx[idx], qx)) < rs && fabsf(__fsub_rn(ky[idx], qy)) < rs)) continue;
                    const int held = kpmp[idx];                                      // :89-91, initial state
                    if (held != -1) {
                        const int obs = held >= 0 ? P.mpObs[mo + held] : (kpobs ? kpobs[idx] : 0);
                        if (obs > 0) continue;
                    }
                    if (ur && ur[idx] > 0) {                                        // :93-98
                        const float er = fabsf(__fsub_rn(qxr, ur[idx]));
                        if (er > rs) continue;
                    }
                    const int dist = hamming256(a0, a1, __ldg(kd + 2 * idx), __ldg(kd + 2 * idx + 1));
                    top4_insert(best, ((uint32_t)dist << 23) | ((uint32_t)p << 5) | (uint32_t)(o & 31));
                    count++;
                }
            }
        }
    }
    P.topk[mo + i] = best;
    P.topkCount[mo + i] = count;
    if (count > 0) {
        const uint32_t k[4] = {best.x, best.y, best.z, best.w};
        uint32_t id[4];
#pragma unroll
        for (int j = 0; j < 4; j++) id[j] = j < count ? (uint32_t)ciw[(k[j] >> 5) & 0x3ffffu] : 0u;
        P.topkIdx[mo + i] = make_uint4(id[0], id[1], id[2], id[3]);
    }
}

// ---- SearchByProjection, phase B: the greedy pass, one warp per frame --------------------------
__global__ void __launch_bounds__(128) k_search_proj(const ProjParams P)
{
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= P.items) return;
    const int n = min(P.f.n[item], P.f.stride), nmp = min(P.mpN[item], P.mpStride);
    const float* kx = P.f.x + (size_t)item * P.f.stride;
    const float* ky = P.f.y + (size_t)item * P.f.stride;
    const int* koct = P.f.octave + (size_t)item * P.f.stride;
    const uint4* kd = reinterpret_cast<const uint4*>(P.f.desc + (size_t)item * P.f.stride * 32);
    const float* ur = P.uRight ? P.uRight + (size_t)item * P.f.stride : nullptr;
    const int* cs = P.cellStart + (size_t)item * (GRID_CELLS + 1);
    const int* ci = P.cellItems + (size_t)item * P.f.stride;
    int* kpmp = P.kpMp + (size_t)item * P.f.stride;
    const int* kpobs = P.kpMpObs ? P.kpMpObs + (size_t)item * P.f.stride : nullptr;
    const size_t mo = (size_t)item * P.mpStride;
    const uint4* md = reinterpret_cast<const uint4*>(P.mpDesc + mo * 32);
    // occupancy of every keypoint (holds a map point with observations -> skipped, :89-91) as one byte in shared
    // memory: the greedy state is read several times per map point and must not cost a global round trip
    extern __shared__ uint8_t s_occ_all[];
    uint8_t* occ = s_occ_all + (size_t)(threadIdx.x >> 5) * ((P.f.stride + 15) & ~15);
    for (int idx = lane; idx < n; idx += 32) {
        const int held = kpmp[idx];
        occ[idx] = held != -1 && (held >= 0 ? P.mpObs[mo + held] : (kpobs ? kpobs[idx] : 0)) > 0;
    }
    __syncwarp();

    int nmatches = 0;
    const bool bFactor = P.th != 1.0f;
    const uint4* topk = P.topk + mo;
    const uint4* topkIdx = P.topkIdx + mo;
    const int* topkCount = P.topkCount + mo;
    // The per-map-point lists do not depend on the greedy state: lane j loads the list of map point base + j
    // (coalesced, one memory latency per 32 map points) and the warp walks them in order by shuffle.
    for (int base = 0; base < nmp; base += 32) {
      const int mine = base + lane;
      int cntL = -1, obsL = 0;
      uint4 kkL = make_uint4(0, 0, 0, 0), idL = kkL;
      if (mine < nmp) { cntL = topkCount[mine]; kkL = topk[mine]; idL = topkIdx[mine]; obsL = P.mpObs[mo + mine]; }
      const int jEnd = min(32, nmp - base);
      for (int j = 0; j < jEnd; j++) {
        const int i = base + j;
        const int cnt = __shfl_sync(0xffffffffu, cntL, j);
        if (cnt <= 0) continue;          // not in view / bad (:56-60), outside the grid, or no candidate (:73)
        const uint4 kk = make_uint4(__shfl_sync(0xffffffffu, kkL.x, j), __shfl_sync(0xffffffffu, kkL.y, j),
                                    __shfl_sync(0xffffffffu, kkL.z, j), __shfl_sync(0xffffffffu, kkL.w, j));
        const uint4 idv = make_uint4(__shfl_sync(0xffffffffu, idL.x, j), __shfl_sync(0xffffffffu, idL.y, j),
                                     __shfl_sync(0xffffffffu, idL.z, j), __shfl_sync(0xffffffffu, idL.w, j));
        const int obsI = __shfl_sync(0xffffffffu, obsL, j);
        Top2 t = {256, INT_MAX, -1, 256, INT_MAX, -1};
        // Fast path (every lane redundantly): the map point's 4 best static candidates in visiting order,
        // minus the keypoints that were taken since (:89-91).  Conclusive when two survive, when the list
        // holds every candidate, or when nothing within TH_HIGH can survive; otherwise the full scan below.
        bool resolved;
        {
            const uint32_t key[4] = {kk.x, kk.y, kk.z, kk.w};
            const uint32_t kid[4] = {idv.x, idv.y, idv.z, idv.w};
            int found = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (found < 2 && j < cnt) {
                    const int dist = (int)(key[j] >> 23), pos = (int)((key[j] >> 5) & 0x3ffffu), o = (int)(key[j] & 31u);
                    const bool taken = occ[kid[j]] != 0;
                    if (!taken && dist < 256) {
                        if (found == 0) { t.b = dist; t.bp = pos; t.ba = o; } else { t.s = dist; t.sp = pos; t.sa = o; }
                        found++;
                    }
                }
            }
            const int d3 = (int)(key[3] >> 23);
            resolved = found == 2 || cnt <= 4 || (found == 1 && t.b > TH_HIGH) || (found == 0 && d3 > TH_HIGH);
        }
        if (!resolved) {
            t = Top2{256, INT_MAX, -1, 256, INT_MAX, -1};
            const int lvl = P.mpLevel[mo + i];
            float r = ((double)P.mpViewCos[mo + i] > 0.998) ? 2.5f : 4.0f;               // :133-139
            if (bFactor) r = __fmul_rn(r, P.th);
            const float rs = __fmul_rn(r, P.scaleFactors[lvl]);
            const float qx = P.mpX[mo + i], qy = P.mpY[mo + i];
            int c0, c1, r0, r1;
            cell_range(P.g, qx, qy, rs, c0, c1, r0, r1);
            const int minLevel = lvl - 1, maxLevel = lvl;
            const bool check = (minLevel > 0) || (maxLevel >= 0);
            const uint4 a0 = __ldg(md + 2 * i), a1 = __ldg(md + 2 * i + 1);
            const float qxr = P.mpXR[mo + i];
            for (int c = c0; c <= c1; c++) {
                const int s = cs[c * GRID_ROWS + r0], e = cs[c * GRID_ROWS + r1 + 1];
                for (int p = s + lane; p < e; p += 32) {
                    const int idx = ci[p];
                    const int o = koct[idx];
                    if (check) {
                        if (o < minLevel) continue;
                        if (maxLevel >= 0 && o > maxLevel) continue;
                    }
                    if (!(fabsf(__fsub_rn(kx[idx], qx)) < rs && fabsf(__fsub_rn(ky[idx], qy)) < rs)) continue;
                    if (occ[idx]) continue;                                              // :89-91
                    if (ur && ur[idx] > 0) {                                            // :93-98
                        const float er = fabsf(__fsub_rn(qxr, ur[idx]));
                        if (er > rs) continue;
                    }
                    const int dist = hamming256(a0, a1, __ldg(kd + 2 * idx), __ldg(kd + 2 * idx + 1));
                    top2_push(t, dist, p, o);
                }
            }
            t = top2_warp_reduce(t);
        }
        if (t.b <= TH_HIGH && !(t.ba == t.sa && (float)t.b > __fmul_rn(P.nnratio, (float)t.s))) {   // :120-127
            if (lane == 0) {
                const int bestIdx = ci[t.bp];
                kpmp[bestIdx] = i;                                                   // :125
                occ[bestIdx] = obsI > 0;
            }
            nmatches++;
        }
        __syncwarp();
      }
    }
    if (lane == 0) P.nmatches[item] = nmatches;
}

// =========================================================================================
// SearchByProjection(CurrentFrame, LastFrame, th, bMono)  (S/ORBmatcher.cc:1332-1474), SURVEY 8(f) N2
// =========================================================================================
struct LastParams {
    FrameDev f;                 // current frame
    const float* uRight;        // items x f.stride or NULL
    GridGeo g;
    const int* cellStart;
    const int* cellItems;
    const int* lastN;           // last frame, items x lastStride
    const uint8_t *hasMp, *outlier;
    const float* wpos;          // x3
    const uint8_t* mpDesc;      // x32
    const int* mpObs;
    const int* lastOct;
    const float* lastAng;
    int lastStride;
    const float *Rcw, *tcw;     // items x 9, items x 3
    float fx, fy, cx, cy, mbf, minX, minY, maxX, maxY;
    int* kpMp;                  // items x f.stride, in/out: index into the last frame's arrays
    const int* kpMpObs;
    const float* scaleFactors;
    int* nmatches;
    uint4 *topk, *topkIdx;      // items x lastStride
    int* topkCount;
    int *histBin, *histIdx;     // items x lastStride scratch
    int items, mode, checkOri;
    float th;
    // kind 1 = the key-frame (relocalisation) overload (:1476-1603): hasMp = "usable map point", no outlier flags,
    // level from MapPoint::PredictScale, any held keypoint is skipped, acceptance threshold orbDist
    int kind, orbDist, nlevels;
    const float *mfMax, *mfMin;  // items x lastStride: raw mfMaxDistance / mfMinDistance
    const float* Ow;             // items x 3
    float logScale;
    // kind 2 = SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) (:294-407): Fuse-style projection and gates on a key
    // frame (int-truncated query bounds, viewing angle), level window [level-1, level], acceptance TH_LOW
    const float* normal;         // items x lastStride x 3
    GridGeo q;                   // query geometry of the key frame
    int maxXi, maxYi;
};

struct LastQuery { float u, v, radius, invzc; int minLevel, maxLevel; bool ok; };

// glibc >= 2.27 logf (ARM optimized-routines): 16-entry {1/c, log c} table + degree-3 polynomial in double, one
// rounding to float.  Bit-identical to libm on every positive finite float (checked exhaustively on the host
// restatement, oracle/orb_matcher_oracle.c:orc_logf); MapPoint::PredictScale depends on it.
__device__ const double d_logf_tab[16][2] = { ORB_B200_LOGF_TABLE };

__device__ __forceinline__ float libm_logf(float x)
{
    uint32_t ix = __float_as_uint(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2 == 0) return -CUDART_INF_F;
        if (ix == 0x7f800000u) return x;
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return CUDART_NAN_F;
        ix = __float_as_uint(__fmul_rn(x, 0x1p23f));
        ix -= 23u << 23;
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u);
    const int k = (int)tmp >> 23;
    const double z = (double)__uint_as_float(ix - (tmp & 0xff800000u));
    const double r = __dsub_rn(__dmul_rn(z, d_logf_tab[i][0]), 1.0);
    const double y0 = __dadd_rn(d_logf_tab[i][1], __dmul_rn((double)k, ORB_B200_LOGF_LN2));
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(ORB_B200_LOGF_A1, r), ORB_B200_LOGF_A2);
    y = __dadd_rn(__dmul_rn(ORB_B200_LOGF_A0, r2), y);
    y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
    return __double2float_rn(y);
}

// projection of one last-frame / key-frame map point into the current frame (:1360-1390, :1500-1532), float
// arithmetic in source order
__device__ __forceinline__ LastQuery last_query(const LastParams& P, int item, int i)
{
    LastQuery q;
    q.ok = false;
    const size_t lo = (size_t)item * P.lastStride + i;
    if (!P.hasMp[lo] || (P.outlier && P.outlier[lo])) return q;
    const float* R = P.Rcw + (size_t)item * 9;
    const float* t = P.tcw + (size_t)item * 3;
    const float* X = P.wpos + lo * 3;
    float c3[3];
#pragma unroll
    for (int r = 0; r < 3; r++)
        c3[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[3 * r], X[0]), __fmul_rn(R[3 * r + 1], X[1])), __fmul_rn(R[3 * r + 2], X[2])), t[r]);
    if (P.kind == 2) {
        if (c3[2] < 0.0f) return q;                                      // :327
        q.invzc = __fdiv_rn(1.0f, c3[2]);
        q.u = __fadd_rn(__fmul_rn(P.fx, __fmul_rn(c3[0], q.invzc)), P.cx);
        q.v = __fadd_rn(__fmul_rn(P.fy, __fmul_rn(c3[1], q.invzc)), P.cy);
        if (!(q.u >= P.q.minX && q.u < (float)P.maxXi && q.v >= P.q.minY && q.v < (float)P.maxYi)) return q;     // KeyFrame::IsInImage
        const float* O = P.Ow + (size_t)item * 3;
        double ss = 0.0, dot = 0.0;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const double po = (double)__fsub_rn(X[r], O[r]);
            ss = __dadd_rn(ss, __dmul_rn(po, po));
            dot = __dadd_rn(dot, __dmul_rn(po, (double)P.normal[lo * 3 + r]));
        }
        const float dist = __double2float_rn(__dsqrt_rn(ss));
        const float mx = P.mfMax[lo];
        if (dist < __fmul_rn(0.8f, P.mfMin[lo]) || dist > __fmul_rn(1.2f, mx)) return q;
        if (dot < __dmul_rn(0.5, (double)dist)) return q;
        int level = (int)ceilf(__fdiv_rn(libm_logf(__fdiv_rn(mx, dist)), P.logScale));
        level = max(0, min(level, P.nlevels - 1));
        q.radius = __fmul_rn(P.th, P.scaleFactors[level]);
        q.minLevel = level - 1; q.maxLevel = level;                      // :371 (maxLevel >= 0, so the level test is active)
        q.ok = true;
        return q;
    }
    q.invzc = (float)__ddiv_rn(1.0, (double)c3[2]);
    if (P.kind == 0 && q.invzc < 0) return q;                            // (the key-frame overload has no depth test)
    q.u = __fadd_rn(__fmul_rn(__fmul_rn(P.fx, c3[0]), q.invzc), P.cx);
    q.v = __fadd_rn(__fmul_rn(__fmul_rn(P.fy, c3[1]), q.invzc), P.cy);
    if (q.u < P.minX || q.u > P.maxX || q.v < P.minY || q.v > P.maxY) return q;
    if (P.kind == 1) {
        const float* O = P.Ow + (size_t)item * 3;
        double ss = 0.0;                                                 // cv::norm(x3Dw - Ow): squares summed in double
#pragma unroll
        for (int r = 0; r < 3; r++) { const double po = (double)__fsub_rn(X[r], O[r]); ss = __dadd_rn(ss, __dmul_rn(po, po)); }
        const float dist3D = __double2float_rn(__dsqrt_rn(ss));
        const float mx = P.mfMax[lo];
        if (dist3D < __fmul_rn(0.8f, P.mfMin[lo]) || dist3D > __fmul_rn(1.2f, mx)) return q;    // :1525-1526
        // MapPoint::PredictScale (S/MapPoint.cc:391-400); clamped like the later upstream fix (the reference indexes
        // mvScaleFactors out of range for dist3D in [0.8 mfMin, mfMin))
        int level = (int)ceilf(__fdiv_rn(libm_logf(__fdiv_rn(mx, dist3D)), P.logScale));
        level = max(0, min(level, P.nlevels - 1));
        q.radius = __fmul_rn(P.th, P.scaleFactors[level]);
        q.minLevel = level - 1; q.maxLevel = level + 1;
        q.ok = true;
        return q;
    }
    const int oct = P.lastOct[lo];
    q.radius = __fmul_rn(P.th, P.scaleFactors[oct]);
    if (P.mode == 1) { q.minLevel = oct; q.maxLevel = -1; }              // bForward  (:1393)
    else if (P.mode == 2) { q.minLevel = 0; q.maxLevel = oct; }          // bBackward (:1395)
    else { q.minLevel = oct - 1; q.maxLevel = oct + 1; }                 // (:1397)
    q.ok = true;
    return q;
}

// static candidate test shared by both phases (GetFeaturesInArea level + window tests, stereo check :1414-1420)
__device__ __forceinline__ bool last_candidate(const LastParams& P, const LastQuery& q, int idx, const float* kx, const float* ky,
                                               const int* koct, const float* ur)
{
    const int o = koct[idx];
    if ((q.minLevel > 0) || (q.maxLevel >= 0)) {
        if (o < q.minLevel) return false;
        if (q.maxLevel >= 0 && o > q.maxLevel) return false;
    }
    if (!(fabsf(__fsub_rn(kx[idx], q.u)) < q.radius && fabsf(__fsub_rn(ky[idx], q.v)) < q.radius)) return false;
    if (ur && ur[idx] > 0) {
        const float pr = __fsub_rn(q.u, __fmul_rn(P.mbf, q.invzc));
        if (fabsf(__fsub_rn(pr, ur[idx])) > q.radius) return false;
    }
    return true;
}

// phase A: one thread per last-frame keypoint
__global__ void __launch_bounds__(128) k_last_topk(const LastParams P)
{
    const int item = blockIdx.y;
    const int nl = min(P.lastN[item], P.lastStride);
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= nl) return;
    const size_t lo = (size_t)item * P.lastStride + i;
    uint4 best = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    int count = -1;
    const int* ci = P.cellItems + (size_t)item * P.f.stride;
    const LastQuery q = last_query(P, item, i);
    int c0, c1, r0, r1;
    if (q.ok && cell_range(P.kind == 2 ? P.q : P.g, q.u, q.v, q.radius, c0, c1, r0, r1)) {
        count = 0;
        const float* kx = P.f.x + (size_t)item * P.f.stride;
        const float* ky = P.f.y + (size_t)item * P.f.stride;
        const int* koct = P.f.octave + (size_t)item * P.f.stride;
        const uint4* kd = reinterpret_cast<const uint4*>(P.f.desc + (size_t)item * P.f.stride * 32);
        const float* ur = P.uRight ? P.uRight + (size_t)item * P.f.stride : nullptr;
        const int* cs = P.cellStart + (size_t)item * (GRID_CELLS + 1);
        const int* kpmp = P.kpMp + (size_t)item * P.f.stride;
        const int* kpobs = P.kpMpObs ? P.kpMpObs + (size_t)item * P.f.stride : nullptr;
        const uint4* md = reinterpret_cast<const uint4*>(P.mpDesc + lo * 32);
        const uint4 a0 = __ldg(md), a1 = __ldg(md + 1);
        for (int c = c0; c <= c1; c++) {
            const int s = cs[c * GRID_ROWS + r0], e = cs[c * GRID_ROWS + r1 + 1];
            for (int p = s; p < e; p++) {
                const int idx = ci[p];
                if (!last_candidate(P, q, idx, kx, ky, koct, ur)) continue;
                const int held = kpmp[idx];                                      // initial occupancy (:1409-1411, :1546-1547)
                if (held != -1 && (P.kind != 0 || (held >= 0 ? P.mpObs[(size_t)item * P.lastStride + held] : (kpobs ? kpobs[idx] : 0)) > 0)) continue;
                const int dist = hamming256(a0, a1, __ldg(kd + 2 * idx), __ldg(kd + 2 * idx + 1));
                top4_insert(best, ((uint32_t)dist << 23) | ((uint32_t)p << 5));
                count++;
            }
        }
    }
    P.topk[lo] = best;
    P.topkCount[lo] = count;
    if (count > 0) {
        const uint32_t k[4] = {best.x, best.y, best.z, best.w};
        uint32_t id[4];
#pragma unroll
        for (int j = 0; j < 4; j++) id[j] = j < count ? (uint32_t)ci[(k[j] >> 5) & 0x3ffffu] : 0u;
        P.topkIdx[lo] = make_uint4(id[0], id[1], id[2], id[3]);
    }
}

// phase B: one warp per frame pair, last-frame keypoints in order
__global__ void __launch_bounds__(128) k_search_last(const LastParams P)
{
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= P.items) return;
    const int n = min(P.f.n[item], P.f.stride), nl = min(P.lastN[item], P.lastStride);
    const float* kx = P.f.x + (size_t)item * P.f.stride;
    const float* ky = P.f.y + (size_t)item * P.f.stride;
    const int* koct = P.f.octave + (size_t)item * P.f.stride;
    const float* kang = P.f.angle + (size_t)item * P.f.stride;
    const uint4* kd = reinterpret_cast<const uint4*>(P.f.desc + (size_t)item * P.f.stride * 32);
    const float* ur = P.uRight ? P.uRight + (size_t)item * P.f.stride : nullptr;
    const int* cs = P.cellStart + (size_t)item * (GRID_CELLS + 1);
    const int* ci = P.cellItems + (size_t)item * P.f.stride;
    int* kpmp = P.kpMp + (size_t)item * P.f.stride;
    const int* kpobs = P.kpMpObs ? P.kpMpObs + (size_t)item * P.f.stride : nullptr;
    const size_t lo = (size_t)item * P.lastStride;
    int* hbin = P.histBin + lo;
    int* hidx = P.histIdx + lo;

    extern __shared__ uint8_t s_occ_all[];
    uint8_t* occ = s_occ_all + (size_t)(threadIdx.x >> 5) * ((P.f.stride + 15) & ~15);
    for (int idx = lane; idx < n; idx += 32) {
        const int held = kpmp[idx];
        occ[idx] = held != -1 && (P.kind != 0 || (held >= 0 ? P.mpObs[lo + held] : (kpobs ? kpobs[idx] : 0)) > 0);
    }
    for (int i = lane; i < nl; i += 32) hbin[i] = -1;
    __syncwarp();

    int nmatches = 0;
    const int accept = P.kind == 1 ? P.orbDist : (P.kind == 2 ? TH_LOW : TH_HIGH);
    for (int base = 0; base < nl; base += 32) {
      const int mine = base + lane;
      int cntL = -1, obsL = 0;
      float angL = 0.f;
      uint4 kkL = make_uint4(0, 0, 0, 0), idL = kkL;
      if (mine < nl) { cntL = P.topkCount[lo + mine]; kkL = P.topk[lo + mine]; idL = P.topkIdx[lo + mine]; obsL = P.kind != 0 ? 1 : P.mpObs[lo + mine]; angL = P.lastAng[lo + mine]; }
      const int jEnd = min(32, nl - base);
      for (int j = 0; j < jEnd; j++) {
        const int i = base + j;
        const int cnt = __shfl_sync(0xffffffffu, cntL, j);
        if (cnt <= 0) continue;
        const uint32_t key[4] = {__shfl_sync(0xffffffffu, kkL.x, j), __shfl_sync(0xffffffffu, kkL.y, j),
                                 __shfl_sync(0xffffffffu, kkL.z, j), __shfl_sync(0xffffffffu, kkL.w, j)};
        const uint32_t kid[4] = {__shfl_sync(0xffffffffu, idL.x, j), __shfl_sync(0xffffffffu, idL.y, j),
                                 __shfl_sync(0xffffffffu, idL.z, j), __shfl_sync(0xffffffffu, idL.w, j)};
        const int obsI = __shfl_sync(0xffffffffu, obsL, j);
        const float angI = __shfl_sync(0xffffffffu, angL, j);
        // only the best candidate matters here (no ratio test): the first entry of the list whose keypoint is free
        int bestDist = 256, bestIdx = -1;
        bool found = false;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            if (!found && e < cnt) {
                const int dist = (int)(key[e] >> 23);
                if (!occ[kid[e]] && dist < 256) { bestDist = dist; bestIdx = (int)kid[e]; found = true; }
            }
        }
        const bool resolved = found || cnt <= 4 || (int)(key[3] >> 23) > accept;
        if (!resolved) {                                   // every listed keypoint was taken: rescan all candidates
            const LastQuery q = last_query(P, item, i);
            int c0, c1, r0, r1;
            cell_range(P.kind == 2 ? P.q : P.g, q.u, q.v, q.radius, c0, c1, r0, r1);
            const uint4* md = reinterpret_cast<const uint4*>(P.mpDesc + (lo + i) * 32);
            const uint4 a0 = __ldg(md), a1 = __ldg(md + 1);
            int bd = 256, bp = INT_MAX;
            for (int c = c0; c <= c1; c++) {
                const int s = cs[c * GRID_ROWS + r0], e = cs[c * GRID_ROWS + r1 + 1];
                for (int p = s + lane; p < e; p += 32) {
                    const int idx = ci[p];
                    if (!last_candidate(P, q, idx, kx, ky, koct, ur) || occ[idx]) continue;
                    const int dist = hamming256(a0, a1, __ldg(kd + 2 * idx), __ldg(kd + 2 * idx + 1));
                    if (dist < bd) { bd = dist; bp = p; }
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const int od = __shfl_xor_sync(0xffffffffu, bd, d), op = __shfl_xor_sync(0xffffffffu, bp, d);
                if (key_lt(od, op, bd, bp)) { bd = od; bp = op; }
            }
            bestDist = bd;
            bestIdx = bd < 256 ? ci[bp] : -1;
        }
        if (bestDist <= accept) {                                                 // :1436-1452, :1561-1579
            if (lane == 0) {
                kpmp[bestIdx] = i;
                occ[bestIdx] = obsI > 0;
                if (P.checkOri) {
                    float rot = __fsub_rn(angI, kang[bestIdx]);
                    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                    int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));
                    if (bin == HISTO_LENGTH) bin = 0;
                    hbin[i] = bin; hidx[i] = bestIdx;
                }
            }
            nmatches++;
        }
        __syncwarp();
      }
    }
    if (P.checkOri) {                                                             // :1455-1471
        __syncwarp();
        int sizes = 0;
        for (int i = 0; i < nl; i += 32) {
            const int b = (i + lane < nl) ? hbin[i + lane] : -1;
            for (int q = 0; q < HISTO_LENGTH; q++) {
                const unsigned m = __ballot_sync(0xffffffffu, b == q);
                if (lane == q) sizes += __popc(m);
            }
        }
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int q = 0; q < HISTO_LENGTH; q++) {
            const int s = __shfl_sync(0xffffffffu, sizes, q);
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = q; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = q; }
            else if (s > max3) { max3 = s; ind3 = q; }
        }
        if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
        else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
        int removed = 0;
        for (int i = lane; i < nl; i += 32) {
            const int b = hbin[i];
            if (b >= 0 && b != ind1 && b != ind2 && b != ind3) { kpmp[hidx[i]] = -1; removed++; }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, d);
        nmatches -= removed;
    }
    if (lane == 0) P.nmatches[item] = nmatches;
}

// ---- SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (S/ORBmatcher.cc:161-292) ----------------------------
// Features are compared only inside a shared vocabulary node, and a frame feature sits in exactly one node, so
// the greedy state (vpMapPointMatches[realIdxF]) never crosses a node: one WARP per (item, key-frame node).  The
// warp finds the node in the frame's sorted node list by binary search, then walks the node's key-frame features
// in order; lanes own the node's frame features (the first 32 keep their descriptors in registers), the
// best / second-best pair is the associative top-2 under (distance, list position), and the winner's lane marks
// its feature as taken.  The rotation histogram needs the whole item and runs in k_bow_finish.
struct BowSide {
    const int* n; const uint8_t* desc; const float* angle; const uint8_t* valid;
    const int* nNodes; const uint32_t* nodeId; const int* nodeStart; const uint32_t* feat;
    int stride, nodeStride;
};
struct BowParams {
    BowSide kf, f;
    int* matches;      // mode 0: items x f.stride, key-frame slot or -1;  mode 1: items x kf.stride, slot of key frame 2 or -1
    int* bins;         // same shape, scratch: rotation bin of an accepted match, else -1
    int* occ;          // mode 1: items x f.stride, vbMatched2
    int* nmatches;
    int items, checkOri;
    int mode;          // 0 = SearchByBoW(pKF, F) (:161-292), 1 = SearchByBoW(pKF1, pKF2) (:526-659)
    float nnratio;
};

__global__ void __launch_bounds__(128) k_bow_match(const BowParams P)
{
    const int lane = threadIdx.x & 31, item = blockIdx.y;
    const int a = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (a >= min(P.kf.nNodes[item], P.kf.nodeStride)) return;
    const uint32_t id = P.kf.nodeId[(size_t)item * P.kf.nodeStride + a];
    const uint32_t* fid = P.f.nodeId + (size_t)item * P.f.nodeStride;
    int lo = 0, hi = min(P.f.nNodes[item], P.f.nodeStride);
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (fid[mid] < id) lo = mid + 1; else hi = mid; }
    if (lo >= min(P.f.nNodes[item], P.f.nodeStride) || fid[lo] != id) return;
    const int* kst = P.kf.nodeStart + (size_t)item * (P.kf.nodeStride + 1);
    const int* fst = P.f.nodeStart + (size_t)item * (P.f.nodeStride + 1);
    const int ks = kst[a], ke = kst[a + 1], fs = fst[lo], fe = fst[lo + 1];
    const uint32_t* kfeat = P.kf.feat + (size_t)item * P.kf.stride;
    const uint32_t* ffeat = P.f.feat + (size_t)item * P.f.stride;
    const uint4* kd = reinterpret_cast<const uint4*>(P.kf.desc + (size_t)item * P.kf.stride * 32);
    const uint4* fd = reinterpret_cast<const uint4*>(P.f.desc + (size_t)item * P.f.stride * 32);
    const uint8_t* kvalid = P.kf.valid ? P.kf.valid + (size_t)item * P.kf.stride : nullptr;
    const uint8_t* fvalid = P.f.valid ? P.f.valid + (size_t)item * P.f.stride : nullptr;
    const int outStride = P.mode ? P.kf.stride : P.f.stride;
    int* matches = P.matches + (size_t)item * outStride;
    int* bins = P.bins + (size_t)item * outStride;
    volatile int* occ = P.mode ? P.occ + (size_t)item * P.f.stride : matches;      // mode 0: a frame keypoint with a match is taken
    const int freeMark = P.mode ? 0 : -1;
    const int nF = fe - fs;
    if (nF <= 0) return;

    // chunk 0 of the frame list lives in registers
    int f0 = lane < nF ? (int)ffeat[fs + lane] : -1;
    if (f0 >= 0 && fvalid && !fvalid[f0]) f0 = -1;                                  // :572-576 (no good map point on side 2)
    uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
    if (f0 >= 0) { r0 = __ldg(fd + 2 * f0); r1 = __ldg(fd + 2 * f0 + 1); }
    bool taken0 = false;

    for (int ik = ks; ik < ke; ik++) {
        const int kidx = (int)kfeat[ik];
        if (kvalid && !kvalid[kidx]) continue;                                  // :193-198 (warp-uniform)
        const uint4 a0 = __ldg(kd + 2 * kidx), a1 = __ldg(kd + 2 * kidx + 1);
        Top2 t = {256, INT_MAX, 0, 256, INT_MAX, 0};
        if (f0 >= 0 && !taken0) top2_push(t, hamming256(a0, a1, r0, r1), lane, f0);
        for (int p = 32 + lane; p < nF; p += 32) {                              // long lists: occupancy from the output array
            const int fi = (int)ffeat[fs + p];
            if (occ[fi] != freeMark || (fvalid && !fvalid[fi])) continue;
            top2_push(t, hamming256(a0, a1, __ldg(fd + 2 * fi), __ldg(fd + 2 * fi + 1)), p, fi);
        }
        t = top2_warp_reduce(t);
        if ((P.mode ? t.b < TH_LOW : t.b <= TH_LOW) && (float)t.b < __fmul_rn(P.nnratio, (float)t.s)) {     // :230-232, :601-603
            if (t.bp == lane) taken0 = true;
            if (lane == 0) {
                const int slot = P.mode ? kidx : t.ba;
                matches[slot] = P.mode ? t.ba : kidx;
                if (P.mode) occ[t.ba] = 1;
                if (P.checkOri) {
                    float rot = __fsub_rn(P.kf.angle[(size_t)item * P.kf.stride + kidx], P.f.angle[(size_t)item * P.f.stride + t.ba]);
                    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                    int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));
                    if (bin == HISTO_LENGTH) bin = 0;
                    bins[slot] = bin;
                }
            }
            if (nF > 32) __syncwarp();
        }
    }
}

// one warp per item: rotation-consistency filter (:273-289) and the match count
__global__ void __launch_bounds__(128) k_bow_finish(const BowParams P)
{
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (item >= P.items) return;
    const BowSide& o = P.mode ? P.kf : P.f;
    const int n = min(o.n[item], o.stride);
    int* matches = P.matches + (size_t)item * o.stride;
    const int* bins = P.bins + (size_t)item * o.stride;
    int count = 0;
    for (int i = lane; i < n; i += 32) count += matches[i] != -1;
    if (P.checkOri) {
        int sizes = 0;
        for (int i = 0; i < n; i += 32) {
            const int b = (i + lane < n) ? bins[i + lane] : -1;
            for (int q = 0; q < HISTO_LENGTH; q++) {
                const unsigned m = __ballot_sync(0xffffffffu, b == q);
                if (lane == q) sizes += __popc(m);
            }
        }
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int q = 0; q < HISTO_LENGTH; q++) {
            const int s = __shfl_sync(0xffffffffu, sizes, q);
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = q; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = q; }
            else if (s > max3) { max3 = s; ind3 = q; }
        }
        if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
        else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
        for (int i = lane; i < n; i += 32) {
            const int b = bins[i];
            if (b >= 0 && b != ind1 && b != ind2 && b != ind3) { matches[i] = -1; count--; }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) count += __shfl_xor_sync(0xffffffffu, count, d);
    if (lane == 0) P.nmatches[item] = count;
}

// ---- SearchForTriangulation (S/ORBmatcher.cc:661-827) -------------------------------------------------------
// No greedy state in this version (vbMatched2 is never set), so every key-frame-1 feature is independent; the
// node-parallel layout of SearchByBoW is kept (one warp per shared node, lanes over the node's key-frame-2
// features).  The sequential rule "dist > bestDist rejects, equality replaces" picks, among the candidates that
// pass the static tests, the smallest distance and of those the LAST in list order: one packed-key warp minimum.
struct TriGeo { const float *x, *y; const int* octave; const float* uRight; const uint8_t* hasMp; };
struct TriParams {
    BowSide k1, k2;
    TriGeo g1, g2;
    const float *F12, *epipole, *scaleFactors2, *levelSigma2;
    int* matches;      // items x k1.stride
    int* bins;
    int onlyStereo, checkOri;
};

__global__ void __launch_bounds__(128) k_tri_match(const TriParams P)
{
    const int lane = threadIdx.x & 31, item = blockIdx.y;
    const int a = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (a >= min(P.k1.nNodes[item], P.k1.nodeStride)) return;
    const uint32_t id = P.k1.nodeId[(size_t)item * P.k1.nodeStride + a];
    const uint32_t* nid2 = P.k2.nodeId + (size_t)item * P.k2.nodeStride;
    const int nn2 = min(P.k2.nNodes[item], P.k2.nodeStride);
    int lo = 0, hi = nn2;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (nid2[mid] < id) lo = mid + 1; else hi = mid; }
    if (lo >= nn2 || nid2[lo] != id) return;
    const int* st1 = P.k1.nodeStart + (size_t)item * (P.k1.nodeStride + 1);
    const int* st2 = P.k2.nodeStart + (size_t)item * (P.k2.nodeStride + 1);
    const int s1 = st1[a], e1 = st1[a + 1], s2 = st2[lo], e2 = st2[lo + 1];
    const size_t o1 = (size_t)item * P.k1.stride, o2 = (size_t)item * P.k2.stride;
    const uint32_t* feat1 = P.k1.feat + o1;
    const uint32_t* feat2 = P.k2.feat + o2;
    const uint4* d1 = reinterpret_cast<const uint4*>(P.k1.desc + o1 * 32);
    const uint4* d2 = reinterpret_cast<const uint4*>(P.k2.desc + o2 * 32);
    const float* F = P.F12 + (size_t)item * 9;
    const float ex = P.epipole[2 * item], ey = P.epipole[2 * item + 1];
    int* matches = P.matches + o1;
    int* bins = P.bins + o1;

    for (int i1 = s1; i1 < e1; i1++) {
        const int idx1 = (int)feat1[i1];
        if (P.g1.hasMp[o1 + idx1]) continue;                                        // :706-708
        const bool stereo1 = P.g1.uRight && P.g1.uRight[o1 + idx1] >= 0;
        if (P.onlyStereo && !stereo1) continue;
        const float x1 = P.g1.x[o1 + idx1], y1 = P.g1.y[o1 + idx1];
        // epipolar line l = x1' F12 (:145-147), float in source order
        const float la = __fadd_rn(__fadd_rn(__fmul_rn(x1, F[0]), __fmul_rn(y1, F[3])), F[6]);
        const float lb = __fadd_rn(__fadd_rn(__fmul_rn(x1, F[1]), __fmul_rn(y1, F[4])), F[7]);
        const float lc = __fadd_rn(__fadd_rn(__fmul_rn(x1, F[2]), __fmul_rn(y1, F[5])), F[8]);
        const float den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
        const uint4 a0 = __ldg(d1 + 2 * idx1), a1 = __ldg(d1 + 2 * idx1 + 1);
        uint32_t best = 0xffffffffu;
        for (int p = lane; p < e2 - s2; p += 32) {
            const int idx2 = (int)feat2[s2 + p];
            if (P.g2.hasMp[o2 + idx2]) continue;
            const bool stereo2 = P.g2.uRight && P.g2.uRight[o2 + idx2] >= 0;
            if (P.onlyStereo && !stereo2) continue;
            const int dist = hamming256(a0, a1, __ldg(d2 + 2 * idx2), __ldg(d2 + 2 * idx2 + 1));
            if (dist > TH_LOW) continue;
            const float x2 = P.g2.x[o2 + idx2], y2 = P.g2.y[o2 + idx2];
            const int oc = P.g2.octave[o2 + idx2];
            if (!stereo1 && !stereo2) {                                              // too close to the epipole (:745-751)
                const float dx = __fsub_rn(ex, x2), dy = __fsub_rn(ey, y2);
                if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.f, P.scaleFactors2[oc])) continue;
            }
            if (den == 0.f) continue;                                                // CheckDistEpipolarLine (:149-158)
            const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, x2), __fmul_rn(lb, y2)), lc);
            const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
            if (!((double)dsqr < __dmul_rn(3.84, (double)P.levelSigma2[oc]))) continue;
            best = min(best, ((uint32_t)dist << 20) | (0xfffffu - (uint32_t)p));
        }
        best = __reduce_min_sync(0xffffffffu, best);
        if (best != 0xffffffffu && lane == 0) {
            const int idx2 = (int)feat2[s2 + (int)(0xfffffu - (best & 0xfffffu))];
            matches[idx1] = idx2;                                                    // :764
            if (P.checkOri) {
                float rot = __fsub_rn(P.k1.angle[o1 + idx1], P.k2.angle[o2 + idx2]);
                if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));
                if (bin == HISTO_LENGTH) bin = 0;
                bins[idx1] = bin;
            }
        }
    }
}

// ---- MapPoint::ComputeDistinctiveDescriptors (S/MapPoint.cc:248-313) ------------------------------------------
// One warp per map point.  Lanes own rows of the n x n distance matrix; a row's median (element (n-1)/2 of the
// sorted row, the row's own 0 included) is found without sorting: the smallest value v with count(d <= v) >= k+1,
// by bisection over 0..256.  n <= 32: the row is computed once (descriptor j broadcast by shuffle) and kept in
// shared memory; larger n: distances are recomputed from the L1-resident descriptors in every bisection step.
__global__ void __launch_bounds__(128) k_distinctive(const int* __restrict__ offsets, const uint8_t* __restrict__ desc, int items,
                                                     int* __restrict__ best, int* __restrict__ bestMedian)
{
    __shared__ uint16_t rowbuf[4][32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int item = blockIdx.x * 4 + w;
    if (item >= items) return;
    const int o = offsets[item], n = offsets[item + 1] - o;
    if (n <= 0) { if (lane == 0) { best[item] = -1; if (bestMedian) bestMedian[item] = 0; } return; }
    const uint4* d = reinterpret_cast<const uint4*>(desc) + 2 * (size_t)o;
    const int k = (n - 1) >> 1;                                        // (int)(0.5*(N-1)) (:298)
    uint32_t key = 0xffffffffu;                                        // (median << 20) | row: first minimum wins
    if (n <= 32) {
        uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
        if (lane < n) { a0 = __ldg(d + 2 * lane); a1 = __ldg(d + 2 * lane + 1); }
        for (int j = 0; j < n; j++) {
            uint4 b0, b1;
            b0.x = __shfl_sync(0xffffffffu, a0.x, j); b0.y = __shfl_sync(0xffffffffu, a0.y, j); b0.z = __shfl_sync(0xffffffffu, a0.z, j); b0.w = __shfl_sync(0xffffffffu, a0.w, j);
            b1.x = __shfl_sync(0xffffffffu, a1.x, j); b1.y = __shfl_sync(0xffffffffu, a1.y, j); b1.z = __shfl_sync(0xffffffffu, a1.z, j); b1.w = __shfl_sync(0xffffffffu, a1.w, j);
            rowbuf[w][lane][j] = (uint16_t)hamming256(a0, a1, b0, b1);
        }
        if (lane < n) {
            int lo = 0, hi = 256;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                int c = 0;
                for (int j = 0; j < n; j++) c += rowbuf[w][lane][j] <= mid;
                if (c >= k + 1) hi = mid; else lo = mid + 1;
            }
            key = ((uint32_t)lo << 20) | (uint32_t)lane;
        }
    } else {
        for (int i = lane; i < n; i += 32) {
            const uint4 a0 = __ldg(d + 2 * i), a1 = __ldg(d + 2 * i + 1);
            int lo = 0, hi = 256;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                int c = 0;
                for (int j = 0; j < n; j++) c += hamming256(a0, a1, __ldg(d + 2 * j), __ldg(d + 2 * j + 1)) <= mid;
                if (c >= k + 1) hi = mid; else lo = mid + 1;
            }
            key = min(key, ((uint32_t)lo << 20) | (uint32_t)i);
        }
    }
    key = __reduce_min_sync(0xffffffffu, key);
    if (lane == 0) { best[item] = (int)(key & 0xfffffu); if (bestMedian) bestMedian[item] = (int)(key >> 20); }
}

// ---- the search inside ORBmatcher::Fuse(pKF, vpMapPoints, th) (S/ORBmatcher.cc:829-975) ------------------------
// Every candidate map point is independent (the keypoints are not consumed; the replace-or-add surgery that follows
// is host code): one thread per map point projects it, applies the frustum / distance / viewing-angle gates,
// predicts the level and scans the key frame's grid cells.  The key frame's grid holds the Frame's assignment
// (float bounds) while its queries use the int-truncated bounds (S/KeyFrame.cc:42, 577-621).
struct FuseParams {
    FrameDev f;                  // the key frame's undistorted keypoints + descriptors
    const float* uRight;         // items x f.stride or NULL (monocular)
    GridGeo g, q;                // assignment / query geometry
    int maxXi, maxYi;
    const int* cellStart; const int* cellItems;
    const int* nmp; const uint8_t* valid; const float *wpos, *normal; const uint8_t* mpDesc; const float *mfMax, *mfMin;
    int mpStride;
    const float *Rcw, *tcw, *Ow;
    float fx, fy, cx, cy, bf, th, logScale;
    const float *scaleFactors, *invLevelSigma2;
    int nlevels;
    int *bestIdx, *bestDist;
    // mode 0: Fuse(pKF, vpMapPoints, th); 1: Fuse(pKF, Scw, ...) (:979-1104, no reprojection-error gates);
    // 2: a SearchBySim3 leg (:1106-1330): second similarity (R2, t2), dist3D = |camera point|, no angle gate, TH_HIGH
    int mode;
    const float *R2, *t2;        // items x 9, items x 3 (mode 2)
};

__global__ void __launch_bounds__(128) k_fuse_search(const FuseParams P)
{
    const int item = blockIdx.y;
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= min(P.nmp[item], P.mpStride)) return;
    const size_t lo = (size_t)item * P.mpStride + i;
    int bestDist = 256, bestIdx = -1;
    do {
        if (!P.valid[lo]) break;
        const float* R = P.Rcw + (size_t)item * 9;
        const float* t = P.tcw + (size_t)item * 3;
        const float* O = P.Ow + (size_t)item * 3;
        const float* X = P.wpos + lo * 3;
        float c3[3];
#pragma unroll
        for (int r = 0; r < 3; r++)
            c3[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[3 * r], X[0]), __fmul_rn(R[3 * r + 1], X[1])), __fmul_rn(R[3 * r + 2], X[2])), t[r]);
        if (P.mode == 2) {                                                          // p3Dc2 = sR21*p3Dc1 + t21 (:1157)
            const float* S = P.R2 + (size_t)item * 9;
            const float* s2 = P.t2 + (size_t)item * 3;
            float d3[3];
#pragma unroll
            for (int r = 0; r < 3; r++)
                d3[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(S[3 * r], c3[0]), __fmul_rn(S[3 * r + 1], c3[1])), __fmul_rn(S[3 * r + 2], c3[2])), s2[r]);
            c3[0] = d3[0]; c3[1] = d3[1]; c3[2] = d3[2];
        }
        if (c3[2] < 0.0f) break;                                                    // :853
        const float invz = __fdiv_rn(1.0f, c3[2]);
        const float u = __fadd_rn(__fmul_rn(P.fx, __fmul_rn(c3[0], invz)), P.cx);
        const float v = __fadd_rn(__fmul_rn(P.fy, __fmul_rn(c3[1], invz)), P.cy);
        if (!(u >= P.q.minX && u < (float)P.maxXi && v >= P.q.minY && v < (float)P.maxYi)) break;     // KeyFrame::IsInImage
        const float ur = __fsub_rn(u, __fmul_rn(P.bf, invz));
        double ss = 0.0, dot = 0.0;                                                 // cv::norm, Mat::dot: double accumulation
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const double po = (double)(P.mode == 2 ? c3[r] : __fsub_rn(X[r], O[r]));
            ss = __dadd_rn(ss, __dmul_rn(po, po));
            if (P.mode != 2) dot = __dadd_rn(dot, __dmul_rn(po, (double)P.normal[lo * 3 + r]));
        }
        const float dist3D = __double2float_rn(__dsqrt_rn(ss));
        const float mx = P.mfMax[lo];
        if (dist3D < __fmul_rn(0.8f, P.mfMin[lo]) || dist3D > __fmul_rn(1.2f, mx)) break;
        if (P.mode != 2 && dot < __dmul_rn(0.5, (double)dist3D)) break;              // viewing angle (:880)
        int level = (int)ceilf(__fdiv_rn(libm_logf(__fdiv_rn(mx, dist3D)), P.logScale));
        level = max(0, min(level, P.nlevels - 1));
        const float radius = __fmul_rn(P.th, P.scaleFactors[level]);
        int c0, c1, r0, r1;
        if (!cell_range(P.q, u, v, radius, c0, c1, r0, r1)) break;
        const float* kx = P.f.x + (size_t)item * P.f.stride;
        const float* ky = P.f.y + (size_t)item * P.f.stride;
        const int* koct = P.f.octave + (size_t)item * P.f.stride;
        const float* kur = P.uRight ? P.uRight + (size_t)item * P.f.stride : nullptr;
        const uint4* kd = reinterpret_cast<const uint4*>(P.f.desc + (size_t)item * P.f.stride * 32);
        const int* cs = P.cellStart + (size_t)item * (GRID_CELLS + 1);
        const int* ci = P.cellItems + (size_t)item * P.f.stride;
        const uint4* md = reinterpret_cast<const uint4*>(P.mpDesc + lo * 32);
        const uint4 a0 = __ldg(md), a1 = __ldg(md + 1);
        for (int c = c0; c <= c1; c++) {
            const int s = cs[c * GRID_ROWS + r0], e = cs[c * GRID_ROWS + r1 + 1];
            for (int p = s; p < e; p++) {
                const int idx = ci[p];
                if (!(fabsf(__fsub_rn(kx[idx], u)) < radius && fabsf(__fsub_rn(ky[idx], v)) < radius)) continue;
                const int kl = koct[idx];
                if (kl < level - 1 || kl > level) continue;                          // :905
                const float ex = __fsub_rn(u, kx[idx]), ey = __fsub_rn(v, ky[idx]);
                float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                const float kr = kur ? kur[idx] : -1.f;
                if (P.mode != 0) {
                    // no reprojection-error gate in these overloads
                } else if (kr >= 0) {
                    const float er = __fsub_rn(ur, kr);
                    e2 = __fadd_rn(e2, __fmul_rn(er, er));
                    if ((double)__fmul_rn(e2, P.invLevelSigma2[kl]) > 7.8) continue;
                } else if ((double)__fmul_rn(e2, P.invLevelSigma2[kl]) > 5.99) continue;
                const int dist = hamming256(a0, a1, __ldg(kd + 2 * idx), __ldg(kd + 2 * idx + 1));
                if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
            }
        }
    } while (false);
    P.bestIdx[lo] = bestDist <= (P.mode == 2 ? TH_HIGH : TH_LOW) ? bestIdx : -1;
    if (P.bestDist) P.bestDist[lo] = bestDist;
}

// ---- DBoW2 transform (Frame::ComputeBoW, S/Frame.cc:520-527; TemplatedVocabulary.h:1133-1266) -------------------
// Phase 1, one thread per descriptor: descend the vocabulary tree (child with the smallest Hamming distance, first on
// ties) to a leaf, remembering the node passed at level L - levelsup.  The tree (a few MB to ~35 MB of node
// descriptors) stays L2-resident.  Phase 2, one CTA per frame: two shared-memory bitonic sorts of (id << 32 | feature)
// give the std::map orders of the BowVector (by word) and the FeatureVector (by node); a word seen c times gets its
// weight added c times as addWeight does, and the L1 norm is accumulated in ascending word order by one thread,
// because the reference's double additions are order dependent.
struct VocDev { const int* childStart; const int* children; const uint4* desc; const int* wordId; const double* weight; int nNodes, L; };

__global__ void __launch_bounds__(128) k_bow_descend(const VocDev V, const int* __restrict__ n, const uint8_t* __restrict__ desc, int stride,
                                                     int levelsup, int* __restrict__ leafOf, int* __restrict__ nodeOf)
{
    const int item = blockIdx.y, f = blockIdx.x * 128 + threadIdx.x;
    if (f >= min(n[item], stride)) return;
    const uint4* d = reinterpret_cast<const uint4*>(desc + ((size_t)item * stride + f) * 32);
    const uint4 a0 = __ldg(d), a1 = __ldg(d + 1);
    const int nidLevel = V.L - levelsup;
    int node = 0, level = 0, nid = 0;
    int cs = V.childStart[0], ce = V.childStart[1];
    do {
        ++level;
        int best = V.children[cs];
        int bestD = hamming256(a0, a1, __ldg(V.desc + 2 * best), __ldg(V.desc + 2 * best + 1));
        for (int c = cs + 1; c < ce; c++) {
            const int id = V.children[c];
            const int dd = hamming256(a0, a1, __ldg(V.desc + 2 * id), __ldg(V.desc + 2 * id + 1));
            if (dd < bestD) { bestD = dd; best = id; }
        }
        node = best;
        if (level == nidLevel) nid = node;
        cs = V.childStart[node]; ce = V.childStart[node + 1];
    } while (ce > cs && level < 64);
    const size_t o = (size_t)item * stride + f;
    leafOf[o] = V.weight[node] > 0 ? node : -1;                        // stopped words (weight 0) drop out (:1164)
    nodeOf[o] = nid;
}

__device__ __forceinline__ void block_bitonic_sort(unsigned long long* key, int P)
{
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = key[i], b = key[ixj];
                    if (((i & k) == 0) == (a > b)) { key[i] = b; key[ixj] = a; }
                }
            }
            __syncthreads();
        }
}

// exclusive rank of every run start among the first `nvalid` sorted keys; returns the number of runs (to all threads)
__device__ __forceinline__ int block_run_ranks(const unsigned long long* key, int nvalid, int* rank, int* scratch)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const int per = (nvalid + nt - 1) / nt, beg = min(tid * per, nvalid), end = min(beg + per, nvalid);
    int local = 0;
    for (int p = beg; p < end; p++) local += (p == 0 || (key[p] >> 32) != (key[p - 1] >> 32));
    scratch[tid] = local;
    __syncthreads();
    if (tid == 0) { int acc = 0; for (int t = 0; t < nt; t++) { const int v = scratch[t]; scratch[t] = acc; acc += v; } scratch[nt] = acc; }
    __syncthreads();
    int r = scratch[tid];
    for (int p = beg; p < end; p++) {
        const bool start = (p == 0 || (key[p] >> 32) != (key[p - 1] >> 32));
        rank[p] = start ? r : -1;
        r += start;
    }
    __syncthreads();
    return scratch[nt];
}

__global__ void __launch_bounds__(256) k_bow_assemble(const VocDev V, const int* __restrict__ n, int stride, int P,
                                                      const int* __restrict__ leafOf, const int* __restrict__ nodeOf,
                                                      int* __restrict__ bowN, uint32_t* __restrict__ bowWord, double* __restrict__ bowValue,
                                                      int* __restrict__ fvN, uint32_t* __restrict__ fvNode, int* __restrict__ fvStart,
                                                      uint32_t* __restrict__ fvFeat)
{
    extern __shared__ __align__(16) unsigned char bow_smem[];
    unsigned long long* key = reinterpret_cast<unsigned long long*>(bow_smem);       // P
    int* rank = reinterpret_cast<int*>(key + P);                                       // P
    int* scratch = rank + P;                                                           // blockDim.x + 1
    __shared__ int sValid;
    __shared__ double sNorm;
    const int item = blockIdx.x, tid = threadIdx.x;
    const int nf = min(n[item], stride);
    const size_t o = (size_t)item * stride;
    if (tid == 0) sValid = 0;
    __syncthreads();

    // ---- BowVector: sort by (word, feature)
    int mine = 0;
    for (int i = tid; i < P; i += blockDim.x) {
        unsigned long long k = ~0ull;
        if (i < nf && leafOf[o + i] >= 0) { k = ((unsigned long long)(unsigned)V.wordId[leafOf[o + i]] << 32) | (unsigned)i; mine++; }
        key[i] = k;
    }
    atomicAdd(&sValid, mine);
    __syncthreads();
    const int nvalid = sValid;
    block_bitonic_sort(key, P);
    const int nb = block_run_ranks(key, nvalid, rank, scratch);
    for (int p = tid; p < nvalid; p += blockDim.x) {
        if (rank[p] < 0) continue;
        int c = 1;
        while (p + c < nvalid && (key[p + c] >> 32) == (key[p] >> 32)) c++;
        const double w = V.weight[leafOf[o + (unsigned)(key[p] & 0xffffffffu)]];
        double v = w;
        for (int q = 1; q < c; q++) v = __dadd_rn(v, w);                                // addWeight, once per occurrence
        bowWord[o + rank[p]] = (uint32_t)(key[p] >> 32);
        bowValue[o + rank[p]] = v;
    }
    __syncthreads();
    if (tid == 0) {                                                                    // BowVector::normalize(L1), in map order
        double norm = 0.0;
        for (int q = 0; q < nb; q++) norm = __dadd_rn(norm, fabs(bowValue[o + q]));
        sNorm = norm;
        bowN[item] = nb;
    }
    __syncthreads();
    if (sNorm > 0.0) for (int q = tid; q < nb; q += blockDim.x) bowValue[o + q] = __ddiv_rn(bowValue[o + q], sNorm);
    __syncthreads();

    // ---- FeatureVector: sort by (node, feature)
    for (int i = tid; i < P; i += blockDim.x)
        key[i] = (i < nf && leafOf[o + i] >= 0) ? (((unsigned long long)(unsigned)nodeOf[o + i] << 32) | (unsigned)i) : ~0ull;
    __syncthreads();
    block_bitonic_sort(key, P);
    const int nn = block_run_ranks(key, nvalid, rank, scratch);
    int* st = fvStart + (size_t)item * (stride + 1);
    for (int p = tid; p < nvalid; p += blockDim.x) {
        fvFeat[o + p] = (uint32_t)(key[p] & 0xffffffffu);
        if (rank[p] >= 0) { fvNode[o + rank[p]] = (uint32_t)(key[p] >> 32); st[rank[p]] = p; }
    }
    if (tid == 0) { st[nn] = nvalid; fvN[item] = nn; }
}

}  // namespace orbb200

// =========================================================================================
// host side
// =========================================================================================
using namespace orbb200;

struct orbb200_matcher {
    int maxItems, maxPoints, device, lastLaunches;
    cudaStream_t stream;
    int *cellStart, *cellItems, *scratchA, *scratchB, *scratchC, *topkCount;
    uint4 *topk, *topkIdx;
    std::vector<void*> allocs;
    // staging for host-pointer calls
    uint8_t* stage; size_t stageBytes;
};

static int m_alloc(orbb200_matcher* m, void** p, size_t bytes)
{
    ORB_CUDA(cudaMalloc(p, std::max<size_t>(bytes, 256)));
    m->allocs.push_back(*p);
    return ORBB200_OK;
}

extern "C" int orbb200_matcher_create(int max_items, int max_points, int device, orbb200_matcher** out)
{
    if (!out || max_items < 1 || max_points < 1) { set_error("invalid matcher parameters"); return ORBB200_EINVAL; }
    *out = nullptr;
    int ndev = orbb200_device_count();
    if (device < 0 || device >= ndev) { set_error("CUDA device %d not available (%d visible)", device, ndev); return ORBB200_ENODEVICE; }
    ORB_CUDA(cudaSetDevice(device));
    orbb200_matcher* m = new orbb200_matcher();
    m->maxItems = max_items; m->maxPoints = max_points; m->device = device; m->lastLaunches = 0;
    m->stage = nullptr; m->stageBytes = 0; m->stream = nullptr;
    int rc = ORBB200_OK;
    const size_t np = (size_t)max_items * max_points;
    if (rc == ORBB200_OK) rc = m_alloc(m, (void**)&m->cellStart, sizeof(int) * (size_t)max_items * (GRID_CELLS + 1));
    if (rc == ORBB200_OK) rc = m_alloc(m, (void**)&m->cellItems, sizeof(int) * np);
    if (rc == ORBB200_OK) rc = m_alloc(m, (void**)&m->scratchA, sizeof(int) * np);
    if (rc == ORBB200_OK) rc = m_alloc(m, (void**)&m->scratchB, sizeof(int) * np);
    if (rc == ORBB200_OK) rc = m_alloc(m, (void**)&m->scratchC, sizeof(int) * np);
    if (rc == ORBB200_OK) rc = m_alloc(m, (void**)&m->topk, sizeof(uint4) * np);
    if (rc == ORBB200_OK) rc = m_alloc(m, (void**)&m->topkCount, sizeof(int) * np);
    if (rc == ORBB200_OK) rc = m_alloc(m, (void**)&m->topkIdx, sizeof(uint4) * np);
    if (rc == ORBB200_OK && cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("cudaStreamCreate failed"); rc = ORBB200_ECUDA;
    }
    if (rc != ORBB200_OK) { orbb200_matcher_destroy(m); return rc; }
    *out = m;
    return ORBB200_OK;
}

extern "C" void orbb200_matcher_destroy(orbb200_matcher* m)
{
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) { cudaStreamSynchronize(m->stream); cudaStreamDestroy(m->stream); }
    for (void* p : m->allocs) cudaFree(p);
    if (m->stage) cudaFree(m->stage);
    delete m;
}
extern "C" void* orbb200_matcher_stream(orbb200_matcher* m) { return m ? (void*)m->stream : nullptr; }
extern "C" int orbb200_matcher_last_launches(const orbb200_matcher* m) { return m ? m->lastLaunches : 0; }
extern "C" int orbb200_matcher_sync(orbb200_matcher* m)
{
    if (!m) { set_error("null handle"); return ORBB200_EINVAL; }
    ORB_CUDA(cudaSetDevice(m->device));
    ORB_CUDA(cudaStreamSynchronize(m->stream));
    return ORBB200_OK;
}

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
static size_t pad(size_t b) { return align_up(b, 256); }

extern "C" int orbb200_descriptor_distance(orbb200_matcher* m, const uint8_t* a, const uint8_t* b, int n, int32_t* dist)
{
    if (!m || !a || !b || !dist || n < 0) { set_error("bad argument"); return ORBB200_EINVAL; }
    if (n == 0) return ORBB200_OK;
    ORB_CUDA(cudaSetDevice(m->device));
    Stager s{m, 0, m->stream};
    int rc = s.reserve(2 * pad((size_t)n * 32) + pad((size_t)n * 4));
    if (rc) return rc;
    const uint8_t *da, *db;
    if ((rc = s.up(a, (size_t)n * 32, &da)) || (rc = s.up(b, (size_t)n * 32, &db))) return rc;
    int* dd = s.out<int>(n);
    k_distance<<<(n + 255) / 256, 256, 0, m->stream>>>(reinterpret_cast<const uint4*>(da), reinterpret_cast<const uint4*>(db), n, dd);
    ORB_CHECK_LAUNCH("k_distance");
    m->lastLaunches = 1;
    ORB_CUDA(cudaMemcpyAsync(dist, dd, sizeof(int) * n, cudaMemcpyDeviceToHost, m->stream));
    ORB_CUDA(cudaStreamSynchronize(m->stream));
    return ORBB200_OK;
}

static GridGeo grid_geo(const float* bounds)
{
    GridGeo g;
    g.minX = bounds[0]; g.minY = bounds[1];                           // Frame::mnMinX/Y (S/Frame.cc:561-589)
    g.invW = (float)GRID_COLS / (bounds[2] - bounds[0]);               // mfGridElementWidthInv (S/Frame.cc:317-318)
    g.invH = (float)GRID_ROWS / (bounds[3] - bounds[1]);
    return g;
}

static int check_view(const orbb200_matcher* m, int items, int stride, const char* what)
{
    if (items < 1 || items > m->maxItems) { set_error("%s: items %d outside 1..%d", what, items, m->maxItems); return ORBB200_EINVAL; }
    if (stride < 1 || stride > m->maxPoints) { set_error("%s: stride %d outside 1..%d", what, stride, m->maxPoints); return ORBB200_EINVAL; }
    if (stride >= (1 << 18)) { set_error("%s: more than 262143 points per item", what); return ORBB200_EINVAL; }
    return ORBB200_OK;
}

static int upload_frame(Stager& s, const orbb200_frame_view* v, int items, FrameDev* d, bool needAngle)
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
static size_t frame_bytes(const orbb200_frame_view* v, int items)
{
    const size_t np = (size_t)items * v->stride;
    return pad(items * 4) + 4 * pad(np * 4) + pad(np * 32);
}
static FrameDev as_dev(const orbb200_frame_view* v)
{
    FrameDev d;
    d.n = v->n; d.x = v->x; d.y = v->y; d.octave = v->octave; d.angle = v->angle; d.desc = v->desc; d.stride = v->stride;
    return d;
}

extern "C" int orbb200_search_for_initialization(orbb200_matcher* m, int items, const orbb200_frame_view* f1,
                                                 const orbb200_frame_view* f2, const float* bounds, float nnratio,
                                                 int check_orientation, int window_size, float* prev_matched,
                                                 int32_t* matches12, int32_t* nmatches, int on_device)
{
    if (!m || !f1 || !f2 || !prev_matched || !matches12 || !nmatches) { set_error("null argument"); return ORBB200_EINVAL; }
    if (!f1->n || !f1->octave || !f1->desc || !f2->n || !f2->x || !f2->y || !f2->octave || !f2->desc ||
        (check_orientation && (!f1->angle || !f2->angle))) { set_error("incomplete frame view"); return ORBB200_EINVAL; }
    int rc;
    if ((rc = check_view(m, items, f1->stride, "f1")) || (rc = check_view(m, items, f2->stride, "f2"))) return rc;
    if (!bounds || !(bounds[2] > bounds[0]) || !(bounds[3] > bounds[1])) { set_error("bad image bounds"); return ORBB200_EINVAL; }
    ORB_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = m->stream;
    InitParams P;
    memset(&P, 0, sizeof(P));
    const size_t np1 = (size_t)items * f1->stride;
    float* dPrev; int *dM12, *dN;
    Stager s{m, 0, st};
    if (on_device) {
        P.f1 = as_dev(f1); P.f2 = as_dev(f2);
        dPrev = prev_matched; dM12 = matches12; dN = nmatches;
    } else {
        if ((rc = s.reserve(frame_bytes(f1, items) + frame_bytes(f2, items) + pad(np1 * 8) + pad(np1 * 4) + pad(items * 4)))) return rc;
        if ((rc = upload_frame(s, f1, items, &P.f1, check_orientation != 0)) || (rc = upload_frame(s, f2, items, &P.f2, check_orientation != 0))) return rc;
        const float* dp;
        if ((rc = s.up(prev_matched, np1 * 2, &dp))) return rc;
        dPrev = const_cast<float*>(dp);
        dM12 = s.out<int>(np1); dN = s.out<int>(items);
    }
    P.g = grid_geo(bounds);
    P.cellStart = m->cellStart; P.cellItems = m->cellItems;
    P.matches21 = m->scratchB; P.histBin = m->scratchC;
    P.prevMatched = dPrev; P.matches12 = dM12; P.nmatches = dN;
    P.topk = m->topk; P.topkCount = m->topkCount; P.topkIdx = m->topkIdx;
    P.items = items; P.window = window_size; P.checkOri = check_orientation; P.nnratio = nnratio;
    // scratch strides follow the views
    k_build_grid<<<items, 256, 0, st>>>(P.f2, P.g, m->cellStart, m->cellItems);
    ORB_CHECK_LAUNCH("k_build_grid");
    k_init_topk<<<dim3((f1->stride + 127) / 128, items), 128, 0, st>>>(P);
    ORB_CHECK_LAUNCH("k_init_topk");
    {
        const size_t sm = 4 * sizeof(uint16_t) * (size_t)((f2->stride + 7) & ~7);
        if (sm > 200 * 1024) { set_error("more than %d keypoints per frame", 25 * 1024); return ORBB200_EINVAL; }
        if (sm > 48 * 1024) ORB_CUDA(cudaFuncSetAttribute(k_search_init, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_search_init<<<(items + 3) / 4, 128, sm, st>>>(P);
    }
    ORB_CHECK_LAUNCH("k_search_init");
    m->lastLaunches = 3;
    if (!on_device) {
        ORB_CUDA(cudaMemcpyAsync(prev_matched, dPrev, np1 * 8, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaMemcpyAsync(matches12, dM12, np1 * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaMemcpyAsync(nmatches, dN, (size_t)items * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaStreamSynchronize(st));
    }
    return ORBB200_OK;
}

extern "C" int orbb200_search_by_projection(orbb200_matcher* m, int items, const orbb200_frame_view* f,
                                            const float* u_right, const orbb200_mappoint_view* mp, int32_t* kp_mp,
                                            const int32_t* kp_mp_obs, const float* scale_factors, int nlevels,
                                            const float* bounds, float nnratio, float th, int32_t* nmatches, int on_device)
{
    if (!m || !f || !mp || !kp_mp || !scale_factors || !nmatches) { set_error("null argument"); return ORBB200_EINVAL; }
    if (!f->n || !f->x || !f->y || !f->octave || !f->desc || !mp->n || !mp->in_view || !mp->bad || !mp->proj_x ||
        !mp->proj_y || !mp->proj_xr || !mp->level || !mp->view_cos || !mp->desc || !mp->obs) { set_error("incomplete view"); return ORBB200_EINVAL; }
    int rc;
    if ((rc = check_view(m, items, f->stride, "frame")) || (rc = check_view(m, items, mp->stride, "map points"))) return rc;
    if (nlevels > 32) { set_error("more than 32 pyramid levels"); return ORBB200_EINVAL; }
    if (nlevels < 1 || !bounds || !(bounds[2] > bounds[0]) || !(bounds[3] > bounds[1])) { set_error("bad geometry"); return ORBB200_EINVAL; }
    ORB_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = m->stream;
    ProjParams P;
    memset(&P, 0, sizeof(P));
    const size_t np = (size_t)items * f->stride, nm = (size_t)items * mp->stride;
    Stager s{m, 0, st};
    int* dN;
    if (on_device) {
        P.f = as_dev(f); P.uRight = u_right;
        P.mpN = mp->n; P.mpInView = mp->in_view; P.mpBad = mp->bad; P.mpX = mp->proj_x; P.mpY = mp->proj_y; P.mpXR = mp->proj_xr;
        P.mpLevel = mp->level; P.mpViewCos = mp->view_cos; P.mpDesc = mp->desc; P.mpObs = mp->obs;
        P.kpMp = kp_mp; P.kpMpObs = kp_mp_obs; P.scaleFactors = scale_factors; dN = nmatches;
    } else {
        const size_t bytes = frame_bytes(f, items) + 3 * pad(np * 4) + pad(items * 4) + 2 * pad(nm) + 6 * pad(nm * 4) + pad(nm * 32) +
                             pad((size_t)nlevels * 4) + pad(items * 4);
        if ((rc = s.reserve(bytes))) return rc;
        if ((rc = upload_frame(s, f, items, &P.f, false))) return rc;
        const int* kpmp;
        if ((rc = s.up(u_right, np, &P.uRight)) || (rc = s.up(kp_mp, np, &kpmp)) || (rc = s.up(kp_mp_obs, np, &P.kpMpObs)) ||
            (rc = s.up(mp->n, items, &P.mpN)) || (rc = s.up(mp->in_view, nm, &P.mpInView)) || (rc = s.up(mp->bad, nm, &P.mpBad)) ||
            (rc = s.up(mp->proj_x, nm, &P.mpX)) || (rc = s.up(mp->proj_y, nm, &P.mpY)) || (rc = s.up(mp->proj_xr, nm, &P.mpXR)) ||
            (rc = s.up(mp->level, nm, &P.mpLevel)) || (rc = s.up(mp->view_cos, nm, &P.mpViewCos)) ||
            (rc = s.up(mp->desc, nm * 32, &P.mpDesc)) || (rc = s.up(mp->obs, nm, &P.mpObs)) ||
            (rc = s.up(scale_factors, (size_t)nlevels, &P.scaleFactors))) return rc;
        P.kpMp = const_cast<int*>(kpmp);
        dN = s.out<int>(items);
    }
    P.mpStride = mp->stride; P.g = grid_geo(bounds); P.cellStart = m->cellStart; P.cellItems = m->cellItems;
    P.nlevels = nlevels; P.nmatches = dN; P.items = items; P.nnratio = nnratio; P.th = th;
    P.topk = m->topk; P.topkCount = m->topkCount; P.topkIdx = m->topkIdx;
    k_build_grid<<<items, 256, 0, st>>>(P.f, P.g, m->cellStart, m->cellItems);
    ORB_CHECK_LAUNCH("k_build_grid");
    k_proj_topk<<<dim3((mp->stride + 127) / 128, items), 128, 0, st>>>(P);
    ORB_CHECK_LAUNCH("k_proj_topk");
    {
        const size_t sm = 4 * (size_t)((f->stride + 15) & ~15);
        if (sm > 200 * 1024) { set_error("more than %d keypoints per frame", 50 * 1024); return ORBB200_EINVAL; }
        if (sm > 48 * 1024) ORB_CUDA(cudaFuncSetAttribute(k_search_proj, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_search_proj<<<(items + 3) / 4, 128, sm, st>>>(P);
    }
    ORB_CHECK_LAUNCH("k_search_proj");
    m->lastLaunches = 3;
    if (!on_device) {
        ORB_CUDA(cudaMemcpyAsync(kp_mp, P.kpMp, np * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaMemcpyAsync(nmatches, dN, (size_t)items * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaStreamSynchronize(st));
    }
    return ORBB200_OK;
}

extern "C" int orbb200_search_by_projection_last_frame(orbb200_matcher* m, int items, const orbb200_frame_view* cur, const float* u_right,
                                                       const orbb200_lastframe_view* last, const float* Rcw, const float* tcw,
                                                       const float* K, float mbf, int32_t* kp_mp, const int32_t* kp_mp_obs,
                                                       const float* scale_factors, int nlevels, const float* bounds, float th,
                                                       int mode, int check_orientation, int32_t* nmatches, int on_device)
{
    if (!m || !cur || !last || !Rcw || !tcw || !K || !kp_mp || !scale_factors || !nmatches) { set_error("null argument"); return ORBB200_EINVAL; }
    if (!cur->n || !cur->x || !cur->y || !cur->octave || !cur->desc || (check_orientation && !cur->angle) || !last->n || !last->has_mp ||
        !last->outlier || !last->world_pos || !last->mp_desc || !last->mp_obs || !last->octave || !last->angle) { set_error("incomplete view"); return ORBB200_EINVAL; }
    int rc;
    if ((rc = check_view(m, items, cur->stride, "current frame")) || (rc = check_view(m, items, last->stride, "last frame"))) return rc;
    if (nlevels < 1 || nlevels > 32 || !bounds || !(bounds[2] > bounds[0]) || !(bounds[3] > bounds[1]) || mode < 0 || mode > 2) { set_error("bad geometry"); return ORBB200_EINVAL; }
    ORB_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = m->stream;
    LastParams P;
    memset(&P, 0, sizeof(P));
    const size_t np = (size_t)items * cur->stride, nl = (size_t)items * last->stride;
    Stager s{m, 0, st};
    int* dN;
    if (on_device) {
        P.f = as_dev(cur); P.uRight = u_right;
        P.lastN = last->n; P.hasMp = last->has_mp; P.outlier = last->outlier; P.wpos = last->world_pos; P.mpDesc = last->mp_desc;
        P.mpObs = last->mp_obs; P.lastOct = last->octave; P.lastAng = last->angle;
        P.Rcw = Rcw; P.tcw = tcw; P.kpMp = kp_mp; P.kpMpObs = kp_mp_obs; P.scaleFactors = scale_factors; dN = nmatches;
    } else {
        const size_t bytes = frame_bytes(cur, items) + 3 * pad(np * 4) + 2 * pad(items * 4) + 2 * pad(nl) + pad(nl * 12) + pad(nl * 32) +
                             3 * pad(nl * 4) + pad((size_t)items * 36) + pad((size_t)items * 12) + pad((size_t)nlevels * 4);
        if ((rc = s.reserve(bytes))) return rc;
        if ((rc = upload_frame(s, cur, items, &P.f, true))) return rc;
        const int* kpmp;
        if ((rc = s.up(u_right, np, &P.uRight)) || (rc = s.up(kp_mp, np, &kpmp)) || (rc = s.up(kp_mp_obs, np, &P.kpMpObs)) ||
            (rc = s.up(last->n, items, &P.lastN)) || (rc = s.up(last->has_mp, nl, &P.hasMp)) || (rc = s.up(last->outlier, nl, &P.outlier)) ||
            (rc = s.up(last->world_pos, nl * 3, &P.wpos)) || (rc = s.up(last->mp_desc, nl * 32, &P.mpDesc)) ||
            (rc = s.up(last->mp_obs, nl, &P.mpObs)) || (rc = s.up(last->octave, nl, &P.lastOct)) || (rc = s.up(last->angle, nl, &P.lastAng)) ||
            (rc = s.up(Rcw, (size_t)items * 9, &P.Rcw)) || (rc = s.up(tcw, (size_t)items * 3, &P.tcw)) ||
            (rc = s.up(scale_factors, (size_t)nlevels, &P.scaleFactors))) return rc;
        P.kpMp = const_cast<int*>(kpmp);
        dN = s.out<int>(items);
    }
    P.lastStride = last->stride; P.g = grid_geo(bounds); P.cellStart = m->cellStart; P.cellItems = m->cellItems;
    P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3]; P.mbf = mbf;
    P.minX = bounds[0]; P.minY = bounds[1]; P.maxX = bounds[2]; P.maxY = bounds[3];
    P.nmatches = dN; P.items = items; P.mode = mode; P.checkOri = check_orientation; P.th = th;
    P.topk = m->topk; P.topkCount = m->topkCount; P.topkIdx = m->topkIdx; P.histBin = m->scratchA; P.histIdx = m->scratchB;
    k_build_grid<<<items, 256, 0, st>>>(P.f, P.g, m->cellStart, m->cellItems);
    ORB_CHECK_LAUNCH("k_build_grid");
    k_last_topk<<<dim3((last->stride + 127) / 128, items), 128, 0, st>>>(P);
    ORB_CHECK_LAUNCH("k_last_topk");
    {
        const size_t sm = 4 * (size_t)((cur->stride + 15) & ~15);
        if (sm > 200 * 1024) { set_error("more than %d keypoints per frame", 50 * 1024); return ORBB200_EINVAL; }
        if (sm > 48 * 1024) ORB_CUDA(cudaFuncSetAttribute(k_search_last, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_search_last<<<(items + 3) / 4, 128, sm, st>>>(P);
    }
    ORB_CHECK_LAUNCH("k_search_last");
    m->lastLaunches = 3;
    if (!on_device) {
        ORB_CUDA(cudaMemcpyAsync(kp_mp, P.kpMp, np * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaMemcpyAsync(nmatches, dN, (size_t)items * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaStreamSynchronize(st));
    }
    return ORBB200_OK;
}

extern "C" int orbb200_search_by_projection_keyframe(orbb200_matcher* m, int items, const orbb200_frame_view* cur,
                                                     const orbb200_keyframe_view* kf, const float* Rcw, const float* tcw,
                                                     const float* Ow, const float* K, int32_t* kp_mp, const float* scale_factors,
                                                     int nlevels, float log_scale_factor, const float* bounds, float th,
                                                     int orb_dist, int check_orientation, int32_t* nmatches, int on_device)
{
    if (!m || !cur || !kf || !Rcw || !tcw || !Ow || !K || !kp_mp || !scale_factors || !nmatches) { set_error("null argument"); return ORBB200_EINVAL; }
    if (!cur->n || !cur->x || !cur->y || !cur->octave || !cur->desc || (check_orientation && !cur->angle) || !kf->n || !kf->valid ||
        !kf->world_pos || !kf->mp_desc || !kf->max_distance || !kf->min_distance || (check_orientation && !kf->angle)) { set_error("incomplete view"); return ORBB200_EINVAL; }
    int rc;
    if ((rc = check_view(m, items, cur->stride, "current frame")) || (rc = check_view(m, items, kf->stride, "key frame"))) return rc;
    if (nlevels < 1 || nlevels > 32 || !bounds || !(bounds[2] > bounds[0]) || !(bounds[3] > bounds[1]) || orb_dist < 0 || orb_dist > 255 ||
        !(log_scale_factor > 0.f)) { set_error("bad geometry or threshold"); return ORBB200_EINVAL; }
    ORB_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = m->stream;
    LastParams P;
    memset(&P, 0, sizeof(P));
    const size_t np = (size_t)items * cur->stride, nl = (size_t)items * kf->stride;
    Stager s{m, 0, st};
    int* dN;
    if (on_device) {
        P.f = as_dev(cur);
        P.lastN = kf->n; P.hasMp = kf->valid; P.wpos = kf->world_pos; P.mpDesc = kf->mp_desc; P.mfMax = kf->max_distance;
        P.mfMin = kf->min_distance; P.lastAng = kf->angle;
        P.Rcw = Rcw; P.tcw = tcw; P.Ow = Ow; P.kpMp = kp_mp; P.scaleFactors = scale_factors; dN = nmatches;
    } else {
        const size_t bytes = frame_bytes(cur, items) + pad(np * 4) + 2 * pad(items * 4) + pad(nl) + pad(nl * 12) + pad(nl * 32) +
                             3 * pad(nl * 4) + pad((size_t)items * 36) + 2 * pad((size_t)items * 12) + pad((size_t)nlevels * 4);
        if ((rc = s.reserve(bytes))) return rc;
        if ((rc = upload_frame(s, cur, items, &P.f, true))) return rc;
        const int* kpmp;
        if ((rc = s.up(kp_mp, np, &kpmp)) || (rc = s.up(kf->n, items, &P.lastN)) || (rc = s.up(kf->valid, nl, &P.hasMp)) ||
            (rc = s.up(kf->world_pos, nl * 3, &P.wpos)) || (rc = s.up(kf->mp_desc, nl * 32, &P.mpDesc)) ||
            (rc = s.up(kf->max_distance, nl, &P.mfMax)) || (rc = s.up(kf->min_distance, nl, &P.mfMin)) ||
            (rc = s.up(kf->angle, kf->angle ? nl : 0, &P.lastAng)) ||
            (rc = s.up(Rcw, (size_t)items * 9, &P.Rcw)) || (rc = s.up(tcw, (size_t)items * 3, &P.tcw)) || (rc = s.up(Ow, (size_t)items * 3, &P.Ow)) ||
            (rc = s.up(scale_factors, (size_t)nlevels, &P.scaleFactors))) return rc;
        P.kpMp = const_cast<int*>(kpmp);
        dN = s.out<int>(items);
    }
    if (!P.lastAng) P.lastAng = P.mfMax;      // never read for a decision when check_orientation is off
    P.kind = 1; P.orbDist = orb_dist; P.nlevels = nlevels; P.logScale = log_scale_factor;
    P.lastStride = kf->stride; P.g = grid_geo(bounds); P.cellStart = m->cellStart; P.cellItems = m->cellItems;
    P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3];
    P.minX = bounds[0]; P.minY = bounds[1]; P.maxX = bounds[2]; P.maxY = bounds[3];
    P.nmatches = dN; P.items = items; P.mode = 0; P.checkOri = check_orientation; P.th = th;
    P.topk = m->topk; P.topkCount = m->topkCount; P.topkIdx = m->topkIdx; P.histBin = m->scratchA; P.histIdx = m->scratchB;
    k_build_grid<<<items, 256, 0, st>>>(P.f, P.g, m->cellStart, m->cellItems);
    ORB_CHECK_LAUNCH("k_build_grid");
    k_last_topk<<<dim3((kf->stride + 127) / 128, items), 128, 0, st>>>(P);
    ORB_CHECK_LAUNCH("k_last_topk");
    {
        const size_t sm = 4 * (size_t)((cur->stride + 15) & ~15);
        if (sm > 200 * 1024) { set_error("more than %d keypoints per frame", 50 * 1024); return ORBB200_EINVAL; }
        if (sm > 48 * 1024) ORB_CUDA(cudaFuncSetAttribute(k_search_last, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_search_last<<<(items + 3) / 4, 128, sm, st>>>(P);
    }
    ORB_CHECK_LAUNCH("k_search_last");
    m->lastLaunches = 3;
    if (!on_device) {
        ORB_CUDA(cudaMemcpyAsync(kp_mp, P.kpMp, np * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaMemcpyAsync(nmatches, dN, (size_t)items * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaStreamSynchronize(st));
    }
    return ORBB200_OK;
}

static int upload_bow_side(Stager& s, const orbb200_bow_view* v, int items, BowSide* d)
{
    int rc;
    const size_t np = (size_t)items * v->stride, nn = (size_t)items * v->node_stride;
    if ((rc = s.up(v->n, items, &d->n)) || (rc = s.up(v->desc, np * 32, &d->desc)) || (rc = s.up(v->angle, v->angle ? np : 0, &d->angle)) ||
        (rc = s.up(v->valid, v->valid ? np : 0, &d->valid)) || (rc = s.up(v->n_nodes, items, &d->nNodes)) ||
        (rc = s.up(v->node_id, nn, &d->nodeId)) || (rc = s.up(v->node_start, nn + items, &d->nodeStart)) || (rc = s.up(v->feat, np, &d->feat))) return rc;
    return ORBB200_OK;
}
static size_t bow_side_bytes(const orbb200_bow_view* v, int items)
{
    const size_t np = (size_t)items * v->stride, nn = (size_t)items * v->node_stride;
    return 2 * pad((size_t)items * 4) + pad(np * 32) + pad(np * 4) + pad(np) + pad(nn * 4) + pad((nn + items) * 4) + pad(np * 4);
}

static int bow_search(orbb200_matcher* m, int items, const orbb200_bow_view* kf, const orbb200_bow_view* f, float nnratio,
                      int check_orientation, int32_t* matches, int32_t* nmatches, int on_device, int mode)
{
    if (!m || !kf || !f || !matches || !nmatches) { set_error("null argument"); return ORBB200_EINVAL; }
    for (const orbb200_bow_view* v : {kf, f})
        if (!v->n || !v->desc || !v->n_nodes || !v->node_id || !v->node_start || !v->feat || (check_orientation && !v->angle) || v->node_stride < 1) {
            set_error("incomplete view"); return ORBB200_EINVAL;
        }
    int rc;
    if ((rc = check_view(m, items, kf->stride, "key frame")) || (rc = check_view(m, items, f->stride, "frame"))) return rc;
    ORB_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = m->stream;
    BowParams P;
    memset(&P, 0, sizeof(P));
    const size_t nf = (size_t)items * (mode ? kf->stride : f->stride);       // output entries
    Stager s{m, 0, st};
    int* dN;
    if (on_device) {
        auto side = [](const orbb200_bow_view* v) {
            BowSide d; d.n = v->n; d.desc = v->desc; d.angle = v->angle; d.valid = v->valid; d.nNodes = v->n_nodes; d.nodeId = v->node_id;
            d.nodeStart = v->node_start; d.feat = v->feat; d.stride = v->stride; d.nodeStride = v->node_stride; return d; };
        P.kf = side(kf); P.f = side(f); P.matches = matches; dN = nmatches;
    } else {
        if ((rc = s.reserve(bow_side_bytes(kf, items) + bow_side_bytes(f, items) + pad(nf * 4) + pad((size_t)items * 4)))) return rc;
        if ((rc = upload_bow_side(s, kf, items, &P.kf)) || (rc = upload_bow_side(s, f, items, &P.f))) return rc;
        P.kf.stride = kf->stride; P.kf.nodeStride = kf->node_stride; P.f.stride = f->stride; P.f.nodeStride = f->node_stride;
        P.matches = s.out<int>(nf);
        dN = s.out<int>(items);
    }
    P.bins = m->scratchA; P.occ = m->scratchB; P.nmatches = dN; P.items = items; P.checkOri = check_orientation; P.nnratio = nnratio;
    P.mode = mode;
    if (mode) ORB_CUDA(cudaMemsetAsync(P.occ, 0, (size_t)items * f->stride * 4, st));
    ORB_CUDA(cudaMemsetAsync(P.matches, 0xff, nf * 4, st));
    if (check_orientation) ORB_CUDA(cudaMemsetAsync(P.bins, 0xff, nf * 4, st));
    k_bow_match<<<dim3((kf->node_stride + 3) / 4, items), 128, 0, st>>>(P);
    ORB_CHECK_LAUNCH("k_bow_match");
    k_bow_finish<<<(items + 3) / 4, 128, 0, st>>>(P);
    ORB_CHECK_LAUNCH("k_bow_finish");
    m->lastLaunches = 2;
    if (!on_device) {
        ORB_CUDA(cudaMemcpyAsync(matches, P.matches, nf * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaMemcpyAsync(nmatches, dN, (size_t)items * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaStreamSynchronize(st));
    }
    return ORBB200_OK;
}

extern "C" int orbb200_search_by_bow(orbb200_matcher* m, int items, const orbb200_bow_view* kf, const orbb200_bow_view* f, float nnratio,
                                     int check_orientation, int32_t* matches, int32_t* nmatches, int on_device)
{
    return bow_search(m, items, kf, f, nnratio, check_orientation, matches, nmatches, on_device, 0);
}

extern "C" int orbb200_search_by_bow_keyframes(orbb200_matcher* m, int items, const orbb200_bow_view* kf1, const orbb200_bow_view* kf2,
                                               float nnratio, int check_orientation, int32_t* matches12, int32_t* nmatches, int on_device)
{
    return bow_search(m, items, kf1, kf2, nnratio, check_orientation, matches12, nmatches, on_device, 1);
}

static int upload_tri_geo(Stager& s, const orbb200_tri_view* v, size_t np, TriGeo* d)
{
    int rc;
    if ((rc = s.up(v->x, np, &d->x)) || (rc = s.up(v->y, np, &d->y)) || (rc = s.up(v->octave, v->octave ? np : 0, &d->octave)) ||
        (rc = s.up(v->u_right, v->u_right ? np : 0, &d->uRight)) || (rc = s.up(v->has_mp, np, &d->hasMp))) return rc;
    return ORBB200_OK;
}

extern "C" int orbb200_search_for_triangulation(orbb200_matcher* m, int items, const orbb200_bow_view* kf1, const orbb200_tri_view* g1,
                                                const orbb200_bow_view* kf2, const orbb200_tri_view* g2, const float* F12,
                                                const float* epipole, const float* scale_factors2, const float* level_sigma2_2,
                                                int nlevels, int only_stereo, int check_orientation, int32_t* matches12,
                                                int32_t* nmatches, int on_device)
{
    if (!m || !kf1 || !kf2 || !g1 || !g2 || !F12 || !epipole || !scale_factors2 || !level_sigma2_2 || !matches12 || !nmatches) { set_error("null argument"); return ORBB200_EINVAL; }
    for (const orbb200_bow_view* v : {kf1, kf2})
        if (!v->n || !v->desc || !v->n_nodes || !v->node_id || !v->node_start || !v->feat || (check_orientation && !v->angle) || v->node_stride < 1) {
            set_error("incomplete view"); return ORBB200_EINVAL;
        }
    if (!g1->x || !g1->y || !g1->has_mp || !g2->x || !g2->y || !g2->octave || !g2->has_mp || nlevels < 1 || nlevels > 32) { set_error("incomplete geometry"); return ORBB200_EINVAL; }
    int rc;
    if ((rc = check_view(m, items, kf1->stride, "key frame 1")) || (rc = check_view(m, items, kf2->stride, "key frame 2"))) return rc;
    if (kf2->stride >= (1 << 20)) { set_error("more than 1048575 features in key frame 2"); return ORBB200_EINVAL; }
    ORB_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = m->stream;
    TriParams P;
    memset(&P, 0, sizeof(P));
    const size_t n1 = (size_t)items * kf1->stride, n2 = (size_t)items * kf2->stride;
    Stager s{m, 0, st};
    int* dN;
    if (on_device) {
        auto side = [](const orbb200_bow_view* v) {
            BowSide d; d.n = v->n; d.desc = v->desc; d.angle = v->angle; d.valid = nullptr; d.nNodes = v->n_nodes; d.nodeId = v->node_id;
            d.nodeStart = v->node_start; d.feat = v->feat; d.stride = v->stride; d.nodeStride = v->node_stride; return d; };
        auto geo = [](const orbb200_tri_view* v) { TriGeo d; d.x = v->x; d.y = v->y; d.octave = v->octave; d.uRight = v->u_right; d.hasMp = v->has_mp; return d; };
        P.k1 = side(kf1); P.k2 = side(kf2); P.g1 = geo(g1); P.g2 = geo(g2);
        P.F12 = F12; P.epipole = epipole; P.scaleFactors2 = scale_factors2; P.levelSigma2 = level_sigma2_2;
        P.matches = matches12; dN = nmatches;
    } else {
        const size_t bytes = bow_side_bytes(kf1, items) + bow_side_bytes(kf2, items) + 4 * pad(n1 * 4) + 4 * pad(n2 * 4) + pad(n1) + pad(n2) +
                             pad((size_t)items * 36) + pad((size_t)items * 8) + 2 * pad((size_t)nlevels * 4) + pad(n1 * 4) + pad((size_t)items * 4);
        if ((rc = s.reserve(bytes))) return rc;
        if ((rc = upload_bow_side(s, kf1, items, &P.k1)) || (rc = upload_bow_side(s, kf2, items, &P.k2)) ||
            (rc = upload_tri_geo(s, g1, n1, &P.g1)) || (rc = upload_tri_geo(s, g2, n2, &P.g2)) ||
            (rc = s.up(F12, (size_t)items * 9, &P.F12)) || (rc = s.up(epipole, (size_t)items * 2, &P.epipole)) ||
            (rc = s.up(scale_factors2, (size_t)nlevels, &P.scaleFactors2)) || (rc = s.up(level_sigma2_2, (size_t)nlevels, &P.levelSigma2))) return rc;
        P.k1.stride = kf1->stride; P.k1.nodeStride = kf1->node_stride; P.k2.stride = kf2->stride; P.k2.nodeStride = kf2->node_stride;
        P.matches = s.out<int>(n1);
        dN = s.out<int>(items);
    }
    P.bins = m->scratchA; P.onlyStereo = only_stereo; P.checkOri = check_orientation;
    ORB_CUDA(cudaMemsetAsync(P.matches, 0xff, n1 * 4, st));
    if (check_orientation) ORB_CUDA(cudaMemsetAsync(P.bins, 0xff, n1 * 4, st));
    k_tri_match<<<dim3((kf1->node_stride + 3) / 4, items), 128, 0, st>>>(P);
    ORB_CHECK_LAUNCH("k_tri_match");
    BowParams B;
    memset(&B, 0, sizeof(B));
    B.kf = P.k1; B.f = P.k2; B.matches = P.matches; B.bins = P.bins; B.nmatches = dN; B.items = items; B.checkOri = check_orientation; B.mode = 1;
    k_bow_finish<<<(items + 3) / 4, 128, 0, st>>>(B);
    ORB_CHECK_LAUNCH("k_bow_finish");
    m->lastLaunches = 2;
    if (!on_device) {
        ORB_CUDA(cudaMemcpyAsync(matches12, P.matches, n1 * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaMemcpyAsync(nmatches, dN, (size_t)items * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaStreamSynchronize(st));
    }
    return ORBB200_OK;
}

extern "C" int orbb200_distinctive_descriptors(orbb200_matcher* m, int items, const int32_t* offsets, const uint8_t* descriptors,
                                               int total, int32_t* best, int32_t* best_median, int on_device)
{
    if (!m || !offsets || !descriptors || !best || items < 1 || total < 0) { set_error("bad argument"); return ORBB200_EINVAL; }
    ORB_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = m->stream;
    const int32_t* dOff = offsets; const uint8_t* dDesc = descriptors; int *dBest = best, *dMed = best_median;
    Stager s{m, 0, st};
    if (!on_device) {
        if (offsets[0] != 0 || offsets[items] != total) { set_error("offsets must run from 0 to total"); return ORBB200_EINVAL; }
        for (int i = 0; i < items; i++)
            if (offsets[i + 1] < offsets[i] || offsets[i + 1] - offsets[i] >= (1 << 20)) { set_error("offsets must ascend (at most 1048575 descriptors per map point)"); return ORBB200_EINVAL; }
        int rc;
        if ((rc = s.reserve(pad(((size_t)items + 1) * 4) + pad((size_t)total * 32 + 32) + 2 * pad((size_t)items * 4)))) return rc;
        if ((rc = s.up(offsets, (size_t)items + 1, &dOff)) || (rc = s.up(descriptors, (size_t)total * 32, &dDesc))) return rc;
        dBest = s.out<int>(items); dMed = s.out<int>(items);
    }
    k_distinctive<<<(items + 3) / 4, 128, 0, st>>>(dOff, dDesc, items, dBest, dMed);
    ORB_CHECK_LAUNCH("k_distinctive");
    m->lastLaunches = 1;
    if (!on_device) {
        ORB_CUDA(cudaMemcpyAsync(best, dBest, (size_t)items * 4, cudaMemcpyDeviceToHost, st));
        if (best_median) ORB_CUDA(cudaMemcpyAsync(best_median, dMed, (size_t)items * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaStreamSynchronize(st));
    }
    return ORBB200_OK;
}

extern "C" int orbb200_fuse_search(orbb200_matcher* m, int items, const orbb200_frame_view* kf, const float* u_right,
                                   const orbb200_fusepoints_view* pts, const float* Rcw, const float* tcw, const float* Ow,
                                   const float* K, float bf, const float* scale_factors, const float* inv_level_sigma2, int nlevels,
                                   float log_scale_factor, const float* bounds, float th, int mode, const float* R2, const float* t2,
                                   int32_t* best_idx, int32_t* best_dist, int on_device)
{
    if (!m || !kf || !pts || !Rcw || !tcw || !K || !scale_factors || !inv_level_sigma2 || !best_idx) { set_error("null argument"); return ORBB200_EINVAL; }
    if (mode < 0 || mode > 2 || (mode == 2 ? (!R2 || !t2) : !Ow)) { set_error("mode %d needs %s", mode, mode == 2 ? "R2 and t2" : "Ow"); return ORBB200_EINVAL; }
    if (!kf->n || !kf->x || !kf->y || !kf->octave || !kf->desc || !pts->n || !pts->valid || !pts->world_pos || (mode != 2 && !pts->normal) || !pts->mp_desc ||
        !pts->max_distance || !pts->min_distance) { set_error("incomplete view"); return ORBB200_EINVAL; }
    int rc;
    if ((rc = check_view(m, items, kf->stride, "key frame")) || (rc = check_view(m, items, pts->stride, "map points"))) return rc;
    if (nlevels < 1 || nlevels > 32 || !bounds || !(bounds[2] > bounds[0]) || !(bounds[3] > bounds[1]) || !(log_scale_factor > 0.f)) { set_error("bad geometry"); return ORBB200_EINVAL; }
    ORB_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = m->stream;
    FuseParams P;
    memset(&P, 0, sizeof(P));
    const size_t np = (size_t)items * kf->stride, nl = (size_t)items * pts->stride;
    Stager s{m, 0, st};
    if (on_device) {
        P.f = as_dev(kf); P.uRight = u_right;
        P.nmp = pts->n; P.valid = pts->valid; P.wpos = pts->world_pos; P.normal = pts->normal; P.mpDesc = pts->mp_desc;
        P.mfMax = pts->max_distance; P.mfMin = pts->min_distance;
        P.Rcw = Rcw; P.tcw = tcw; P.Ow = Ow; P.scaleFactors = scale_factors; P.invLevelSigma2 = inv_level_sigma2;
        P.R2 = R2; P.t2 = t2;
        P.bestIdx = best_idx; P.bestDist = best_dist;
    } else {
        const size_t bytes = frame_bytes(kf, items) + pad(np * 4) + pad((size_t)items * 4) + pad(nl) + 2 * pad(nl * 12) + pad(nl * 32) + 2 * pad(nl * 4) +
                             2 * pad((size_t)items * 36) + 3 * pad((size_t)items * 12) + 2 * pad((size_t)nlevels * 4) + 2 * pad(nl * 4);
        if ((rc = s.reserve(bytes))) return rc;
        if ((rc = upload_frame(s, kf, items, &P.f, false))) return rc;
        if ((rc = s.up(mode == 2 ? R2 : nullptr, (size_t)items * 9, &P.R2)) || (rc = s.up(mode == 2 ? t2 : nullptr, (size_t)items * 3, &P.t2))) return rc;
        if ((rc = s.up(u_right, u_right ? np : 0, &P.uRight)) || (rc = s.up(pts->n, items, &P.nmp)) || (rc = s.up(pts->valid, nl, &P.valid)) ||
            (rc = s.up(pts->world_pos, nl * 3, &P.wpos)) || (rc = s.up(mode != 2 ? pts->normal : nullptr, nl * 3, &P.normal)) || (rc = s.up(pts->mp_desc, nl * 32, &P.mpDesc)) ||
            (rc = s.up(pts->max_distance, nl, &P.mfMax)) || (rc = s.up(pts->min_distance, nl, &P.mfMin)) ||
            (rc = s.up(Rcw, (size_t)items * 9, &P.Rcw)) || (rc = s.up(tcw, (size_t)items * 3, &P.tcw)) || (rc = s.up(mode != 2 ? Ow : nullptr, (size_t)items * 3, &P.Ow)) ||
            (rc = s.up(scale_factors, (size_t)nlevels, &P.scaleFactors)) || (rc = s.up(inv_level_sigma2, (size_t)nlevels, &P.invLevelSigma2))) return rc;
        P.bestIdx = s.out<int>(nl);
        P.bestDist = s.out<int>(nl);
    }
    P.mpStride = pts->stride; P.g = grid_geo(bounds); P.q = P.g;
    P.q.minX = (float)(int)bounds[0]; P.q.minY = (float)(int)bounds[1];          // KeyFrame::mnMinX/Y are ints (S/KeyFrame.cc:42)
    P.maxXi = (int)bounds[2]; P.maxYi = (int)bounds[3];
    P.cellStart = m->cellStart; P.cellItems = m->cellItems;
    P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3]; P.bf = bf; P.th = th; P.logScale = log_scale_factor; P.nlevels = nlevels;
    P.mode = mode;
    k_build_grid<<<items, 256, 0, st>>>(P.f, P.g, m->cellStart, m->cellItems);
    ORB_CHECK_LAUNCH("k_build_grid");
    k_fuse_search<<<dim3((pts->stride + 127) / 128, items), 128, 0, st>>>(P);
    ORB_CHECK_LAUNCH("k_fuse_search");
    m->lastLaunches = 2;
    if (!on_device) {
        ORB_CUDA(cudaMemcpyAsync(best_idx, P.bestIdx, nl * 4, cudaMemcpyDeviceToHost, st));
        if (best_dist) ORB_CUDA(cudaMemcpyAsync(best_dist, P.bestDist, nl * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaStreamSynchronize(st));
    }
    return ORBB200_OK;
}

extern "C" int orbb200_search_by_projection_sim3(orbb200_matcher* m, int items, const orbb200_frame_view* kf,
                                                 const orbb200_fusepoints_view* pts, const float* Rcw, const float* tcw, const float* Ow,
                                                 const float* K, const float* scale_factors, int nlevels, float log_scale_factor,
                                                 const float* bounds, int th, int32_t* matched, int32_t* nmatches, int on_device)
{
    if (!m || !kf || !pts || !Rcw || !tcw || !Ow || !K || !scale_factors || !matched || !nmatches) { set_error("null argument"); return ORBB200_EINVAL; }
    if (!kf->n || !kf->x || !kf->y || !kf->octave || !kf->desc || !pts->n || !pts->valid || !pts->world_pos || !pts->normal || !pts->mp_desc ||
        !pts->max_distance || !pts->min_distance) { set_error("incomplete view"); return ORBB200_EINVAL; }
    int rc;
    if ((rc = check_view(m, items, kf->stride, "key frame")) || (rc = check_view(m, items, pts->stride, "map points"))) return rc;
    if (nlevels < 1 || nlevels > 32 || !bounds || !(bounds[2] > bounds[0]) || !(bounds[3] > bounds[1]) || !(log_scale_factor > 0.f)) { set_error("bad geometry"); return ORBB200_EINVAL; }
    ORB_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = m->stream;
    LastParams P;
    memset(&P, 0, sizeof(P));
    const size_t np = (size_t)items * kf->stride, nl = (size_t)items * pts->stride;
    Stager s{m, 0, st};
    int* dN;
    if (on_device) {
        P.f = as_dev(kf);
        P.lastN = pts->n; P.hasMp = pts->valid; P.wpos = pts->world_pos; P.normal = pts->normal; P.mpDesc = pts->mp_desc;
        P.mfMax = pts->max_distance; P.mfMin = pts->min_distance;
        P.Rcw = Rcw; P.tcw = tcw; P.Ow = Ow; P.kpMp = matched; P.scaleFactors = scale_factors; dN = nmatches;
    } else {
        const size_t bytes = frame_bytes(kf, items) + pad(np * 4) + 2 * pad((size_t)items * 4) + pad(nl) + 2 * pad(nl * 12) + pad(nl * 32) +
                             2 * pad(nl * 4) + pad((size_t)items * 36) + 2 * pad((size_t)items * 12) + pad((size_t)nlevels * 4);
        if ((rc = s.reserve(bytes))) return rc;
        if ((rc = upload_frame(s, kf, items, &P.f, false))) return rc;
        const int* kpmp;
        if ((rc = s.up(matched, np, &kpmp)) || (rc = s.up(pts->n, items, &P.lastN)) || (rc = s.up(pts->valid, nl, &P.hasMp)) ||
            (rc = s.up(pts->world_pos, nl * 3, &P.wpos)) || (rc = s.up(pts->normal, nl * 3, &P.normal)) || (rc = s.up(pts->mp_desc, nl * 32, &P.mpDesc)) ||
            (rc = s.up(pts->max_distance, nl, &P.mfMax)) || (rc = s.up(pts->min_distance, nl, &P.mfMin)) ||
            (rc = s.up(Rcw, (size_t)items * 9, &P.Rcw)) || (rc = s.up(tcw, (size_t)items * 3, &P.tcw)) || (rc = s.up(Ow, (size_t)items * 3, &P.Ow)) ||
            (rc = s.up(scale_factors, (size_t)nlevels, &P.scaleFactors))) return rc;
        P.kpMp = const_cast<int*>(kpmp);
        dN = s.out<int>(items);
    }
    P.lastAng = P.mfMax;                       // orientation is not checked in this overload
    P.kind = 2; P.nlevels = nlevels; P.logScale = log_scale_factor;
    P.lastStride = pts->stride; P.g = grid_geo(bounds); P.q = P.g;
    P.q.minX = (float)(int)bounds[0]; P.q.minY = (float)(int)bounds[1]; P.maxXi = (int)bounds[2]; P.maxYi = (int)bounds[3];
    P.cellStart = m->cellStart; P.cellItems = m->cellItems;
    P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3];
    P.nmatches = dN; P.items = items; P.mode = 0; P.checkOri = 0; P.th = (float)th;
    P.topk = m->topk; P.topkCount = m->topkCount; P.topkIdx = m->topkIdx; P.histBin = m->scratchA; P.histIdx = m->scratchB;
    k_build_grid<<<items, 256, 0, st>>>(P.f, P.g, m->cellStart, m->cellItems);
    ORB_CHECK_LAUNCH("k_build_grid");
    k_last_topk<<<dim3((pts->stride + 127) / 128, items), 128, 0, st>>>(P);
    ORB_CHECK_LAUNCH("k_last_topk");
    {
        const size_t sm = 4 * (size_t)((kf->stride + 15) & ~15);
        if (sm > 200 * 1024) { set_error("more than %d keypoints per frame", 50 * 1024); return ORBB200_EINVAL; }
        if (sm > 48 * 1024) ORB_CUDA(cudaFuncSetAttribute(k_search_last, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_search_last<<<(items + 3) / 4, 128, sm, st>>>(P);
    }
    ORB_CHECK_LAUNCH("k_search_last");
    m->lastLaunches = 3;
    if (!on_device) {
        ORB_CUDA(cudaMemcpyAsync(matched, P.kpMp, np * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaMemcpyAsync(nmatches, dN, (size_t)items * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaStreamSynchronize(st));
    }
    return ORBB200_OK;
}

struct orbb200_vocabulary {
    int device;
    VocDev V;
    std::vector<void*> allocs;
};

extern "C" void orbb200_vocabulary_destroy(orbb200_vocabulary* v)
{
    if (!v) return;
    cudaSetDevice(v->device);
    for (void* p : v->allocs) cudaFree(p);
    delete v;
}

extern "C" int orbb200_vocabulary_create(int device, int n_nodes, int levels, const int32_t* child_start, const int32_t* children,
                                         const uint8_t* descriptors, const int32_t* word_id, const double* weight,
                                         orbb200_vocabulary** out)
{
    if (!out || !child_start || !children || !descriptors || !word_id || !weight || n_nodes < 2 || levels < 1 || levels > 32) { set_error("invalid vocabulary"); return ORBB200_EINVAL; }
    *out = nullptr;
    if (child_start[0] != 0 || child_start[1] <= 0) { set_error("the root (node 0) needs children"); return ORBB200_EINVAL; }
    for (int i = 0; i < n_nodes; i++)
        if (child_start[i + 1] < child_start[i]) { set_error("child_start must ascend"); return ORBB200_EINVAL; }
    const int nc = child_start[n_nodes];
    for (int c = 0; c < nc; c++)
        if (children[c] <= 0 || children[c] >= n_nodes) { set_error("child id out of range"); return ORBB200_EINVAL; }
    int ndev = orbb200_device_count();
    if (device < 0 || device >= ndev) { set_error("CUDA device %d not available (%d visible)", device, ndev); return ORBB200_ENODEVICE; }
    ORB_CUDA(cudaSetDevice(device));
    orbb200_vocabulary* v = new orbb200_vocabulary();
    v->device = device;
    auto up = [&](const void* src, size_t bytes, const void** dst) -> int {
        void* p = nullptr;
        if (cudaMalloc(&p, std::max<size_t>(bytes, 256)) != cudaSuccess) { set_error("cudaMalloc failed"); return ORBB200_ECUDA; }
        v->allocs.push_back(p);
        if (cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("cudaMemcpy failed"); return ORBB200_ECUDA; }
        *dst = p;
        return ORBB200_OK;
    };
    int rc;
    if ((rc = up(child_start, sizeof(int32_t) * ((size_t)n_nodes + 1), (const void**)&v->V.childStart)) ||
        (rc = up(children, sizeof(int32_t) * (size_t)std::max(nc, 1), (const void**)&v->V.children)) ||
        (rc = up(descriptors, (size_t)n_nodes * 32, (const void**)&v->V.desc)) ||
        (rc = up(word_id, sizeof(int32_t) * (size_t)n_nodes, (const void**)&v->V.wordId)) ||
        (rc = up(weight, sizeof(double) * (size_t)n_nodes, (const void**)&v->V.weight))) { orbb200_vocabulary_destroy(v); return rc; }
    v->V.nNodes = n_nodes; v->V.L = levels;
    *out = v;
    return ORBB200_OK;
}

extern "C" int orbb200_bow_transform(orbb200_matcher* m, const orbb200_vocabulary* voc, int items, const int32_t* n, const uint8_t* desc,
                                     int stride, int levelsup, int32_t* bow_n, uint32_t* bow_word, double* bow_value, int32_t* fv_n_nodes,
                                     uint32_t* fv_node_id, int32_t* fv_node_start, uint32_t* fv_feat, int on_device)
{
    if (!m || !voc || !n || !desc || !bow_n || !bow_word || !bow_value || !fv_n_nodes || !fv_node_id || !fv_node_start || !fv_feat) { set_error("null argument"); return ORBB200_EINVAL; }
    if (voc->device != m->device) { set_error("vocabulary and matcher live on different devices"); return ORBB200_EINVAL; }
    int rc;
    if ((rc = check_view(m, items, stride, "frame"))) return rc;
    if (stride > 8192) { set_error("more than 8192 features per frame"); return ORBB200_ECAPACITY; }
    ORB_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = m->stream;
    const size_t np = (size_t)items * stride;
    Stager s{m, 0, st};
    const int* dN = n; const uint8_t* dDesc = desc;
    int *dBowN = bow_n, *dFvN = fv_n_nodes, *dFvStart = fv_node_start;
    uint32_t *dBowWord = bow_word, *dFvNode = fv_node_id, *dFvFeat = fv_feat;
    double* dBowValue = bow_value;
    if (!on_device) {
        if ((rc = s.reserve(pad((size_t)items * 4) + pad(np * 32) + 2 * pad((size_t)items * 4) + 3 * pad(np * 4) + pad(np * 8) + pad(((size_t)items * (stride + 1)) * 4)))) return rc;
        if ((rc = s.up(n, items, &dN)) || (rc = s.up(desc, np * 32, &dDesc))) return rc;
        dBowN = s.out<int>(items); dFvN = s.out<int>(items);
        dBowWord = s.out<uint32_t>(np); dFvNode = s.out<uint32_t>(np); dFvFeat = s.out<uint32_t>(np);
        dBowValue = s.out<double>(np);
        dFvStart = s.out<int>((size_t)items * (stride + 1));
    }
    int* leafOf = m->scratchA; int* nodeOf = m->scratchB;
    k_bow_descend<<<dim3((stride + 127) / 128, items), 128, 0, st>>>(voc->V, dN, dDesc, stride, levelsup, leafOf, nodeOf);
    ORB_CHECK_LAUNCH("k_bow_descend");
    int P = 32;
    while (P < stride) P <<= 1;
    const size_t sm = (size_t)P * 12 + 257 * 4 + 16;
    if (sm > 48 * 1024) ORB_CUDA(cudaFuncSetAttribute(k_bow_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_bow_assemble<<<items, 256, sm, st>>>(voc->V, dN, stride, P, leafOf, nodeOf, dBowN, dBowWord, dBowValue, dFvN, dFvNode, dFvStart, dFvFeat);
    ORB_CHECK_LAUNCH("k_bow_assemble");
    m->lastLaunches = 2;
    if (!on_device) {
        ORB_CUDA(cudaMemcpyAsync(bow_n, dBowN, (size_t)items * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaMemcpyAsync(fv_n_nodes, dFvN, (size_t)items * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaMemcpyAsync(bow_word, dBowWord, np * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaMemcpyAsync(bow_value, dBowValue, np * 8, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaMemcpyAsync(fv_node_id, dFvNode, np * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaMemcpyAsync(fv_node_start, dFvStart, (size_t)items * (stride + 1) * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaMemcpyAsync(fv_feat, dFvFeat, np * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaStreamSynchronize(st));
    }
    return ORBB200_OK;
}
