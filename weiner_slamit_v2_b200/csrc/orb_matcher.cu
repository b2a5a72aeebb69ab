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
#include "matcher_common.cuh"

namespace orbb200 {

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
__global__ void __launch_bounds__(256) k_build_grid(FrameDev f, GridGeo g, int* cellStart, int* cellItems, uint4* cellRec)
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
    // The same keypoints once more as 48-byte records IN CSR ORDER: {x, y, index, octave} and the descriptor.  A search
    // that walks a cell range then reads consecutive records instead of chasing index -> four arrays -> descriptor.
    if (!cellRec) return;
    __syncthreads();
    uint4* rec = cellRec + (size_t)item * f.stride * 3;
    const int* oc = f.octave + (size_t)item * f.stride;
    const uint4* de = reinterpret_cast<const uint4*>(f.desc + (size_t)item * f.stride * 32);
    const int ng = cs[GRID_CELLS];
    for (int p = tid; p < ng; p += 256) {
        const int i = ci[p];
        rec[3 * p] = make_uint4(__float_as_uint(kx[i]), __float_as_uint(ky[i]), (uint32_t)i, (uint32_t)oc[i]);
        rec[3 * p + 1] = __ldg(de + 2 * i);
        rec[3 * p + 2] = __ldg(de + 2 * i + 1);
    }
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
    int farDist;              // every second-best at this distance or beyond passes the ratio test against any best <= TH_LOW
};

// ---- SearchForInitialization, phase A: everything that does not depend on the greedy state ----
// One THREAD per two F1 keypoints (queries); the F2 keypoints are streamed through shared memory in CSR order
// (= GetFeaturesInArea visiting order) and every lane tests the same candidate at the same time, so the
// candidate's descriptor is one broadcast load.  The 4 smallest keys (distance, visiting position) and the
// candidate count are kept per query.
//
// The loop is bound by the integer pipes, not by bytes (tools/int_peak.cu: 15 POPC and 63 LOP3/IADD3 lanes per
// clock and SM): a warp-wide 256-bit distance in the plain form holds the POPC pipe for 64 cycles, and every other
// ALU instruction of the loop body costs 2 more.  So
//  (1) distances use the carry-save form (5 POPC + 14 LOP3 instead of 8 + 8);
//  (2) a warp whose queries all see the whole grid (window larger than the image: cell range = all cells, and
//      |x - qx| < r holds for the bounding box of the indexed keypoints, hence for every keypoint because the float
//      subtraction is monotonic) and are all on level 0 skips the per-candidate window tests;
//  (3) on that path only candidates that can matter are evaluated in full.  A match needs best <= TH_LOW and
//      best < ratio * second (:463-465); with farDist = the smallest distance d for which TH_LOW < ratio * d, any
//      second-best at farDist or beyond passes that test whatever its exact value, and any best there fails the first.
//      Candidates at farDist or beyond are therefore interchangeable: the list keeps the 4 smallest keys among the
//      NEAR candidates (distance < farDist), and when fewer than 4 are near the count reported to the greedy pass is
//      the number of near ones, so it knows the list is complete.  The distance over the first 128 bits is a lower
//      bound: a candidate whose half distance reaches farDist is dropped after 3 POPC + 6 LOP3; the survivors (about
//      one in ten on unrelated descriptors) are queued per lane in shared memory and finished in dense batches,
//      because finishing them in place would make every lane pay for the few that survive;
//  (4) elsewhere (windowed searches) the cell-range test of GetFeaturesInArea is two packed adds and one LOP3.
constexpr int TOPK_CHUNK = 256;
struct __align__(16) CandMeta { float x, y; uint32_t cg; int oct; };     // cg = cx | cy << 16 (PosInGrid cell)
constexpr int TOPK_QPT = 2;                        // queries per thread
constexpr int TOPK_QPB = 128 * TOPK_QPT;            // queries per block
constexpr int TOPK_SUB = 256;                       // candidates between two survivor flushes at the latest (positions are stored as bytes)
constexpr int TOPK_LIST = 32;                       // survivor slots per query between flushes

__global__ void __launch_bounds__(128, 8) k_init_topk(const InitParams P)
{
    __shared__ __align__(16) uint4 s_desc[TOPK_CHUNK * 2];
    __shared__ __align__(16) CandMeta s_meta[TOPK_CHUNK];
    __shared__ float s_box[4][4];
    __shared__ int s_oct0;                                       // running number of staged level-0 candidates
    __shared__ uint8_t s_list[TOPK_QPT * TOPK_LIST * 128];       // [query slot][entry][thread]: position of a survivor in its sub-block
    const int item = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n1 = min(P.f1.n[item], P.f1.stride);
    if ((int)blockIdx.x * TOPK_QPB >= n1) return;
    const float* k2x = P.f2.x + (size_t)item * P.f2.stride;
    const float* k2y = P.f2.y + (size_t)item * P.f2.stride;
    const int* oct2 = P.f2.octave + (size_t)item * P.f2.stride;
    const uint4* d2 = reinterpret_cast<const uint4*>(P.f2.desc + (size_t)item * P.f2.stride * 32);
    const int* cs = P.cellStart + (size_t)item * (GRID_CELLS + 1);
    const int* ci = P.cellItems + (size_t)item * P.f2.stride;
    const int ngrid = cs[GRID_CELLS];                          // keypoints that are in the grid at all

    // bounding box of the indexed F2 keypoints
    {
        float bx0 = INFINITY, bx1 = -INFINITY, by0 = INFINITY, by1 = -INFINITY;
        for (int c = tid; c < ngrid; c += 128) {
            const int i2 = ci[c];
            const float x = k2x[i2], y = k2y[i2];
            bx0 = fminf(bx0, x); bx1 = fmaxf(bx1, x); by0 = fminf(by0, y); by1 = fmaxf(by1, y);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            bx0 = fminf(bx0, __shfl_xor_sync(0xffffffffu, bx0, d)); bx1 = fmaxf(bx1, __shfl_xor_sync(0xffffffffu, bx1, d));
            by0 = fminf(by0, __shfl_xor_sync(0xffffffffu, by0, d)); by1 = fmaxf(by1, __shfl_xor_sync(0xffffffffu, by1, d));
        }
        if (lane == 0) { s_box[warp][0] = bx0; s_box[warp][1] = bx1; s_box[warp][2] = by0; s_box[warp][3] = by1; }
        if (tid == 0) s_oct0 = 0;
    }
    __syncthreads();
    const float bx0 = fminf(fminf(s_box[0][0], s_box[1][0]), fminf(s_box[2][0], s_box[3][0]));
    const float bx1 = fmaxf(fmaxf(s_box[0][1], s_box[1][1]), fmaxf(s_box[2][1], s_box[3][1]));
    const float by0 = fminf(fminf(s_box[0][2], s_box[1][2]), fminf(s_box[2][2], s_box[3][2]));
    const float by1 = fmaxf(fmaxf(s_box[0][3], s_box[1][3]), fmaxf(s_box[2][3], s_box[3][3]));

    // the thread's queries: q[u] = block base + u * 128 + tid
    const float r = (float)P.window;
    int qi[TOPK_QPT], level1[TOPK_QPT], octHi[TOPK_QPT], count[TOPK_QPT], near[TOPK_QPT];
    bool active[TOPK_QPT], whole = true, anyActive = false;
    float qx[TOPK_QPT], qy[TOPK_QPT];
    uint32_t loK[TOPK_QPT], hiK[TOPK_QPT];
    uint4 a0[TOPK_QPT], a1[TOPK_QPT], best[TOPK_QPT];
#pragma unroll
    for (int u = 0; u < TOPK_QPT; u++) {
        qi[u] = blockIdx.x * TOPK_QPB + u * 128 + tid;
        active[u] = qi[u] < n1;
        level1[u] = 0; qx[u] = qy[u] = 0.f; count[u] = 0; near[u] = 0;
        a0[u] = a1[u] = make_uint4(0, 0, 0, 0);
        best[u] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        int c0 = 0, c1 = -1, r0 = 0, r1 = -1;
        if (active[u]) {
            level1[u] = P.f1.octave[(size_t)item * P.f1.stride + qi[u]];
            const float* prev = P.prevMatched + ((size_t)item * P.f1.stride + qi[u]) * 2;
            qx[u] = prev[0]; qy[u] = prev[1];
            active[u] = level1[u] <= 0 && cell_range(P.g, qx[u], qy[u], r, c0, c1, r0, r1);
            const uint4* d1 = reinterpret_cast<const uint4*>(P.f1.desc + ((size_t)item * P.f1.stride + qi[u]) * 32);
            a0[u] = __ldg(d1); a1[u] = __ldg(d1 + 1);
        }
        const bool w = level1[u] == 0 && c0 == 0 && c1 == GRID_COLS - 1 && r0 == 0 && r1 == GRID_ROWS - 1 &&
                       fabsf(__fsub_rn(bx0, qx[u])) < r && fabsf(__fsub_rn(bx1, qx[u])) < r &&
                       fabsf(__fsub_rn(by0, qy[u])) < r && fabsf(__fsub_rn(by1, qy[u])) < r;
        whole = whole && (!active[u] || w);
        anyActive = anyActive || active[u];
        // cell-range test on packed 16-bit fields: with cg = cx | cy << 16 (0 <= cx, cy < 64),
        // cg + loK has bit 15 / 31 set iff cx >= c0 / cy >= r0 and cg + hiK has them set iff cx > c1 / cy > r1
        loK[u] = (uint32_t)(0x8000 - c0) | ((uint32_t)(0x8000 - r0) << 16);
        hiK[u] = (uint32_t)(0x7fff - max(c1, 0)) | ((uint32_t)(0x7fff - max(r1, 0)) << 16);
        octHi[u] = level1[u] >= 0 ? level1[u] : INT_MAX;                                   // Frame.cc:468-485
    }
    const bool warpWhole = __all_sync(0xffffffffu, whole);
    const bool warpIdle = !__any_sync(0xffffffffu, anyActive);

    for (int base = 0; base < ngrid; base += TOPK_CHUNK) {
        const int nc = min(TOPK_CHUNK, ngrid - base);
        __syncthreads();
        int lvl0 = 0;
        for (int c = tid; c < nc; c += 128) {
            const int i2 = ci[base + c];
            const float x = k2x[i2], y = k2y[i2];
            const int cx = (int)roundf(__fmul_rn(__fsub_rn(x, P.g.minX), P.g.invW));
            const int cy = (int)roundf(__fmul_rn(__fsub_rn(y, P.g.minY), P.g.invH));
            CandMeta m;
            m.x = x; m.y = y; m.cg = (uint32_t)cx | ((uint32_t)cy << 16); m.oct = oct2[i2];
            lvl0 += m.oct == 0;
            s_meta[c] = m;
            s_desc[2 * c] = __ldg(d2 + 2 * i2);
            s_desc[2 * c + 1] = __ldg(d2 + 2 * i2 + 1);
        }
        if (lvl0) atomicAdd(&s_oct0, lvl0);
        __syncthreads();
        if (warpIdle) continue;
        if (warpWhole) {
            // (explicit shared-space addresses: with generic pointers the compiler rebuilds the shared window base
            // for every candidate and carries 64-bit list pointers)
            const int far = P.farDist;
            for (int sub = 0; sub < nc; sub += TOPK_SUB) {
                const int ns = min(TOPK_SUB, nc - sub);
                const uint32_t sdesc = (uint32_t)__cvta_generic_to_shared(s_desc + 2 * sub);
                const uint32_t soct = (uint32_t)__cvta_generic_to_shared(&s_meta[sub].oct);
                const uint32_t lb0 = (uint32_t)__cvta_generic_to_shared(s_list + tid), lb1 = lb0 + TOPK_LIST * 128;
                uint32_t lp0 = lb0, lp1 = lb1;                                  // next free entry of the lane's two survivor lists
                // finish the queued survivors: full distance, near ones enter the top-4 list
                auto flush_one = [&](int u, uint32_t lbase, uint32_t& lptr) {
                    const int n = (int)(lptr - lbase) >> 7;
                    const int nmax = __reduce_max_sync(0xffffffffu, n);
                    for (int j = 0; j < nmax; j++) {
                        if (j < n) {
                            const int c = sub + (int)lds_u8(lbase + 128 * j);
                            const int dist = hamming256_csa(a0[u], a1[u], s_desc[2 * c], s_desc[2 * c + 1]);
                            if (dist < far) { top4_insert(best[u], ((uint32_t)dist << 20) | (uint32_t)(base + c)); near[u]++; }
                        }
                    }
                    lptr = lbase;
                };
                for (int c4 = 0; c4 < ns; c4 += 4) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int c = c4 + k;
                        if (c < ns && lds_u32(soct + 16 * c) == 0u) {                          // warp-uniform
                            const uint4 b0 = lds_v4(sdesc + 32 * c);
                            if (hamming128_lower(a0[0], b0) < far) { sts_u8(lp0, (uint32_t)c); lp0 += 128; }
                            if (hamming128_lower(a0[1], b0) < far) { sts_u8(lp1, (uint32_t)c); lp1 += 128; }
                        }
                    }
                    // at most 4 entries were added per list since the last check
                    if (__any_sync(0xffffffffu, max(lp0 - lb0, lp1 - lb1) > (uint32_t)((TOPK_LIST - 4) * 128))) {
                        flush_one(0, lb0, lp0); flush_one(1, lb1, lp1);
                    }
                }
                flush_one(0, lb0, lp0); flush_one(1, lb1, lp1);
            }
#pragma unroll
            for (int u = 0; u < TOPK_QPT; u++) count[u] = s_oct0;
        } else {
            for (int c = 0; c < nc; c++) {
                const CandMeta m = s_meta[c];
#pragma unroll
                for (int u = 0; u < TOPK_QPT; u++) {
                    if (!active[u]) continue;
                    if ((((m.cg + loK[u]) & ~(m.cg + hiK[u])) & 0x80008000u) != 0x80008000u) continue;
                    if (m.oct < level1[u] || m.oct > octHi[u]) continue;
                    if (!(fabsf(__fsub_rn(m.x, qx[u])) < r && fabsf(__fsub_rn(m.y, qy[u])) < r)) continue;
                    const int dist = hamming256_csa(a0[u], a1[u], s_desc[2 * c], s_desc[2 * c + 1]);
                    top4_insert(best[u], ((uint32_t)dist << 20) | (uint32_t)(base + c));
                    count[u]++; near[u]++;
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < TOPK_QPT; u++) {
        if (qi[u] >= n1) continue;
        const size_t o = (size_t)item * P.f1.stride + qi[u];
        // fewer than 4 near candidates: the list holds all of them and nothing beyond it can matter (see (3) above)
        const int reported = near[u] >= 4 ? count[u] : near[u];
        P.topk[o] = best[u];
        P.topkCount[o] = active[u] ? reported : -1;
        if (active[u] && reported > 0) {
            const uint32_t k[4] = {best[u].x, best[u].y, best[u].z, best[u].w};
            uint32_t id[4];
#pragma unroll
            for (int j = 0; j < 4; j++) id[j] = j < reported ? (uint32_t)ci[k[j] & 0xfffffu] : 0u;
            P.topkIdx[o] = make_uint4(id[0], id[1], id[2], id[3]);
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
    const int padded = (P.f2.stride + 15) & ~15;
    uint16_t* vmd = s_vmd_all + (size_t)(threadIdx.x >> 5) * padded;
    uint8_t* claim = reinterpret_cast<uint8_t*>(s_vmd_all + (size_t)(blockDim.x >> 5) * padded) + (size_t)(threadIdx.x >> 5) * padded;
    for (int i = lane; i < n2; i += 32) { vmd[i] = 0xffff; m21[i] = -1; claim[i] = 0xff; }
    for (int i = lane; i < n1; i += 32) { m12[i] = -1; hbin[i] = -1; }
    __syncwarp();

    const float r = (float)P.window;
    const uint4* topk = P.topk + (size_t)item * P.f1.stride;
    const uint4* topkIdx = P.topkIdx + (size_t)item * P.f1.stride;
    const int* topkCount = P.topkCount + (size_t)item * P.f1.stride;
    // Lane j owns query base + j and decides it SPECULATIVELY against the current vMatchedDistance, all 32 at once; a
    // decision is final when no earlier query of the batch takes a keypoint the lane looked at (its <= 4 list
    // entries).  The longest clean prefix is committed in parallel, the rest is decided again.
    for (int base = 0; base < n1; base += 32) {
      const int mine = base + lane;
      int cnt = -1;
      uint4 kk = make_uint4(0, 0, 0, 0), idv = kk;
      if (mine < n1) { cnt = topkCount[mine]; kk = topk[mine]; idv = topkIdx[mine]; }
      // cnt <= 0: octave > 0 (:425-427), query outside the grid, or no candidate (:431)
      unsigned todo = __ballot_sync(0xffffffffu, cnt > 0);
      const uint32_t key[4] = {kk.x, kk.y, kk.z, kk.w};
      const int kid[4] = {cnt > 0 ? (int)idv.x : -1, cnt > 1 ? (int)idv.y : -1, cnt > 2 ? (int)idv.z : -1, cnt > 3 ? (int)idv.w : -1};
      while (todo) {
        const int first = __ffs(todo) - 1;
        const bool pending = (todo >> lane) & 1u;
        // Walk the query's 4 best static candidates in visiting order and drop those a previous query already holds
        // at a distance <= ours (:448).  The first two survivors are best / second-best.  That is conclusive when
        // two survive, when the list holds every candidate, or when the best survivor already fails TH_LOW;
        // otherwise the full scan below.
        int tb = INT_MAX, ts = INT_MAX, best2 = -1;
        bool resolved = true;
        if (pending) {
            int found = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (found < 2 && j < cnt) {
                    const int dist = (int)(key[j] >> 20);
                    if (!((int)vmd[kid[j]] <= dist)) {
                        if (found == 0) { tb = dist; best2 = kid[j]; } else ts = dist;
                        found++;
                    }
                }
            }
            // every candidate beyond the list is at least as far as its last entry (d3)
            const int d3 = (int)(key[3] >> 20);
            resolved = found == 2 || cnt <= 4 || (found == 1 && tb > TH_LOW) || (found == 0 && d3 > TH_LOW);
            if (!resolved && found == 1 && (float)tb < __fmul_rn((float)d3, P.nnratio)) {
                ts = d3;                   // the true second-best is >= d3: the ratio test passes either way
                resolved = true;
            }
        }
        bool accept;
        int stop;
        if (__shfl_sync(0xffffffffu, (int)!resolved, first)) {
            // the lowest pending query sees the exact state and its list is inconclusive: the warp scans for it
            const int i1 = base + first;
            const int level1 = oct1[i1];
            Top2 t = Top2{INT_MAX, INT_MAX, -1, INT_MAX, INT_MAX, -1};
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
            if (lane == first) { tb = t.b; ts = t.s; best2 = t.b < INT_MAX ? ci[t.bp] : -1; }
            accept = lane == first && t.b <= TH_LOW && (float)t.b < __fmul_rn((float)t.s, P.nnratio);   // :463-465
            stop = first + 1;
        } else {
            accept = pending && resolved && tb <= TH_LOW && (float)tb < __fmul_rn((float)ts, P.nnratio);   // :463-465
            stop = clean_prefix(claim, lane, accept ? best2 : -1, kid, pending && !resolved);
        }
        if (accept && lane < stop) {          // the committed queries of one round take distinct keypoints
            const int old = m21[best2];
            if (old >= 0) m12[old] = -1;                                             // :467-471
            m12[mine] = best2;
            m21[best2] = mine;
            vmd[best2] = (uint16_t)tb;
            if (P.checkOri) {
                float rot = __fsub_rn(ang1[mine], ang2[best2]);
                if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));           // sic (:417,:482)
                if (bin == HISTO_LENGTH) bin = 0;
                hbin[mine] = bin;
            }
        }
        todo = stop < 32 ? (todo & (0xffffffffu << stop)) : 0u;
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
    const uint4* cellRec;
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
        const float* ur = P.uRight ? P.uRight + (size_t)item * P.f.stride : nullptr;
        const int* cs = P.cellStart + (size_t)item * (GRID_CELLS + 1);
        const int* kpmp = P.kpMp + (size_t)item * P.f.stride;
        const int* kpobs = P.kpMpObs ? P.kpMpObs + (size_t)item * P.f.stride : nullptr;
        const int lvl = P.mpLevel[mo + i];
        float r = ((double)P.mpViewCos[mo + i] > 0.998) ? 2.5f : 4.0f;               // :133-139
        if (P.th != 1.0f) r = __fmul_rn(r, P.th);
        // MapPoint::PredictScale (S/MapPoint.cc:391-400) is unclamped, so the caller's level may lie outside the table:
        // the radius uses the nearest valid entry (the reference reads out of bounds there), minLevel / maxLevel keep the raw value
        const float rs = __fmul_rn(r, P.scaleFactors[min(max(lvl, 0), P.nlevels - 1)]);
        const float qx = P.mpX[mo + i], qy = P.mpY[mo + i], qxr = P.mpXR[mo + i];
        int c0, c1, r0, r1;
        if (cell_range(P.g, qx, qy, rs, c0, c1, r0, r1)) {
            count = 0;
            const int minLevel = lvl - 1, maxLevel = lvl;
            const bool check = (minLevel > 0) || (maxLevel >= 0);
            const uint4* md = reinterpret_cast<const uint4*>(P.mpDesc + (mo + i) * 32);
            const uint4 a0 = __ldg(md), a1 = __ldg(md + 1);
            const uint4* cr = P.cellRec + (size_t)item * P.f.stride * 3;
            for (int c = c0; c <= c1; c++) {
                const int s = cs[c * GRID_ROWS + r0], e = cs[c * GRID_ROWS + r1 + 1];
                for (int p = s; p < e; p++) {
                    const uint4 rec = __ldg(cr + 3 * p);                            // {x, y, index, octave} in CSR order
                    const uint4 b0 = __ldg(cr + 3 * p + 1), b1 = __ldg(cr + 3 * p + 2);
                    const int idx = (int)rec.z, o = (int)rec.w;
                    if (check) {
                        if (o < minLevel) continue;
                        if (maxLevel >= 0 && o > maxLevel) continue;
                    }
                    if (!(fabsf(__fsub_rn(__uint_as_float(rec.x), qx)) < rs && fabsf(__fsub_rn(__uint_as_float(rec.y), qy)) < rs)) continue;
                    const int held = kpmp[idx];                                      // :89-91, initial state
                    if (held != -1) {
                        const int obs = held >= 0 ? P.mpObs[mo + held] : (kpobs ? kpobs[idx] : 0);
                        if (obs > 0) continue;
                    }
                    if (ur && ur[idx] > 0) {                                        // :93-98
                        const float er = fabsf(__fsub_rn(qxr, ur[idx]));
                        if (er > rs) continue;
                    }
                    const int dist = hamming256(a0, a1, b0, b1);
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
    const float* ur = P.uRight ? P.uRight + (size_t)item * P.f.stride : nullptr;
    const int* cs = P.cellStart + (size_t)item * (GRID_CELLS + 1);
    const int* ci = P.cellItems + (size_t)item * P.f.stride;
    const uint4* cr = P.cellRec + (size_t)item * P.f.stride * 3;
    int* kpmp = P.kpMp + (size_t)item * P.f.stride;
    const int* kpobs = P.kpMpObs ? P.kpMpObs + (size_t)item * P.f.stride : nullptr;
    const size_t mo = (size_t)item * P.mpStride;
    const uint4* md = reinterpret_cast<const uint4*>(P.mpDesc + mo * 32);
    // occupancy of every keypoint (holds a map point with observations -> skipped, :89-91) as one byte in shared
    // memory: the greedy state is read several times per map point and must not cost a global round trip
    extern __shared__ uint8_t s_occ_all[];
    uint8_t* occ = s_occ_all + (size_t)(threadIdx.x >> 5) * 2 * ((P.f.stride + 15) & ~15);
    uint8_t* claim = occ + ((P.f.stride + 15) & ~15);
    for (int idx = lane; idx < n; idx += 32) {
        const int held = kpmp[idx];
        occ[idx] = held != -1 && (held >= 0 ? P.mpObs[mo + held] : (kpobs ? kpobs[idx] : 0)) > 0;
        claim[idx] = 0xff;
    }
    __syncwarp();

    int nmatches = 0;
    const bool bFactor = P.th != 1.0f;
    const uint4* topk = P.topk + mo;
    const uint4* topkIdx = P.topkIdx + mo;
    const int* topkCount = P.topkCount + mo;
    // The per-map-point lists do not depend on the greedy state.  Lane j owns map point base + j and decides it
    // SPECULATIVELY against the current occupancy, all 32 at once; a decision is final when no earlier map point of
    // the batch takes a keypoint the lane looked at (its <= 4 list entries).  The longest clean prefix is committed
    // in parallel, the rest is decided again: typically two rounds per 32 map points instead of 32 serial steps.
    for (int base = 0; base < nmp; base += 32) {
      const int mine = base + lane;
      int cnt = -1, obs = 0;
      uint4 kk = make_uint4(0, 0, 0, 0), idv = kk;
      if (mine < nmp) { cnt = topkCount[mine]; kk = topk[mine]; idv = topkIdx[mine]; obs = P.mpObs[mo + mine]; }
      // cnt <= 0: not in view / bad (:56-60), outside the grid, or no candidate (:73)
      unsigned todo = __ballot_sync(0xffffffffu, cnt > 0);
      const uint32_t key[4] = {kk.x, kk.y, kk.z, kk.w};
      const int kid[4] = {cnt > 0 ? (int)idv.x : -1, cnt > 1 ? (int)idv.y : -1, cnt > 2 ? (int)idv.z : -1, cnt > 3 ? (int)idv.w : -1};
      while (todo) {
        const int first = __ffs(todo) - 1;
        const bool pending = (todo >> lane) & 1u;
        // the map point's 4 best static candidates in visiting order, minus the keypoints that were taken since
        // (:89-91).  Conclusive when two survive, when the list holds every candidate, or when nothing within
        // TH_HIGH can survive; otherwise the full scan below.
        Top2 t = {256, INT_MAX, -1, 256, INT_MAX, -1};
        int bestId = -1;
        bool resolved = true;
        if (pending) {
            int found = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (found < 2 && j < cnt) {
                    const int dist = (int)(key[j] >> 23), o = (int)(key[j] & 31u);
                    const bool taken = occ[kid[j]] != 0;
                    if (!taken && dist < 256) {
                        if (found == 0) { t.b = dist; t.ba = o; bestId = kid[j]; } else { t.s = dist; t.sa = o; }
                        found++;
                    }
                }
            }
            const int d3 = (int)(key[3] >> 23);
            resolved = found == 2 || cnt <= 4 || (found == 1 && t.b > TH_HIGH) || (found == 0 && d3 > TH_HIGH);
        }
        if (__shfl_sync(0xffffffffu, (int)!resolved, first)) {
            // the lowest pending map point sees the exact state and its list is inconclusive: the warp scans for it
            const int i = base + first;
            const int obsI = __shfl_sync(0xffffffffu, obs, first);
            t = Top2{256, INT_MAX, -1, 256, INT_MAX, -1};
            const int lvl = P.mpLevel[mo + i];
            float r = ((double)P.mpViewCos[mo + i] > 0.998) ? 2.5f : 4.0f;               // :133-139
            if (bFactor) r = __fmul_rn(r, P.th);
            const float rs = __fmul_rn(r, P.scaleFactors[min(max(lvl, 0), P.nlevels - 1)]);
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
                    const uint4 rec = __ldg(cr + 3 * p);                            // {x, y, index, octave} in CSR order
                    const uint4 b0 = __ldg(cr + 3 * p + 1), b1 = __ldg(cr + 3 * p + 2);
                    const int idx = (int)rec.z, o = (int)rec.w;
                    if (check) {
                        if (o < minLevel) continue;
                        if (maxLevel >= 0 && o > maxLevel) continue;
                    }
                    if (!(fabsf(__fsub_rn(__uint_as_float(rec.x), qx)) < rs && fabsf(__fsub_rn(__uint_as_float(rec.y), qy)) < rs)) continue;
                    if (occ[idx]) continue;                                              // :89-91
                    if (ur && ur[idx] > 0) {                                            // :93-98
                        const float er = fabsf(__fsub_rn(qxr, ur[idx]));
                        if (er > rs) continue;
                    }
                    top2_push(t, hamming256(a0, a1, b0, b1), p, o);
                }
            }
            t = top2_warp_reduce(t);
            if (t.b <= TH_HIGH && !(t.ba == t.sa && (float)t.b > __fmul_rn(P.nnratio, (float)t.s))) {   // :120-127
                if (lane == 0) {
                    const int bestIdx = ci[t.bp];
                    kpmp[bestIdx] = i;                                                   // :125
                    occ[bestIdx] = obsI > 0;
                }
                nmatches++;
            }
            todo &= ~(1u << first);
            __syncwarp();
            continue;
        }
        const bool accept = pending && resolved && t.b <= TH_HIGH &&
                            !(t.ba == t.sa && (float)t.b > __fmul_rn(P.nnratio, (float)t.s));          // :120-127
        const int w = (accept && obs > 0) ? bestId : -1;       // the occupancy this map point would set
        const int stop = clean_prefix(claim, lane, w, kid, pending && !resolved);
        const bool commit = accept && lane < stop;
        nmatches += __popc(__ballot_sync(0xffffffffu, commit));
        // several map points without observations may take the same keypoint; the last one in list order stays (:125)
        const unsigned same = __match_any_sync(0xffffffffu, commit ? bestId : -1 - lane);
        if (commit) {
            if (lane == 31 - __clz(same)) kpmp[bestId] = mine;
            if (obs > 0) occ[bestId] = 1;
        }
        todo = stop < 32 ? (todo & (0xffffffffu << stop)) : 0u;
        __syncwarp();
      }
    }
    if (lane == 0) P.nmatches[item] = nmatches;
}

// ---- SearchByProjection, phase B, frame-wide: one CTA per frame, deterministic reservations ---------------------------------
// The greedy pass above is one dependent chain per frame (10 k map points in list order, 32 at a time): 0.53 of the call's
// 1.05 ms at 512 frames, and no shorter with fewer frames per GPU.  Here every pending map point i posts its index with
// atomicMin on each keypoint of its candidate set C_i -- the keypoints that pass the state-independent tests of :62-98, i.e.
// everything i can ever read (occupancy) or write (its best) -- and a map point whose index survives on all of them is FINAL: no
// earlier pending map point shares a keypoint with it, so the occupancy it sees is the one the sequential loop would show it.
// Final map points decide and commit together, the rest go round again; the lowest pending index is always final, so the
// rounds end, after about as many as map points share a keypoint (5-10 in configs[4]) instead of 313 dependent batches.
// Decisions are the reference's: best / second among the unoccupied candidates in visiting order (from the top-4 list when it
// is conclusive, else a rescan under the current occupancy), ratio test on equal levels, the last accepting map point in list
// order stays on a keypoint that is never occupied (atomicMax).
constexpr int SP2_THREADS = 512;     // (256 threads and four frames per SM measured slower: 0.95 against 0.83 ms)

template <typename F>
__device__ __forceinline__ void proj_for_each_candidate(const ProjParams& P, int item, size_t mo, int i, F&& f)
{
    const int* cs = P.cellStart + (size_t)item * (GRID_CELLS + 1);
    const uint4* cr = P.cellRec + (size_t)item * P.f.stride * 3;
    const float* ur = P.uRight ? P.uRight + (size_t)item * P.f.stride : nullptr;
    const int lvl = P.mpLevel[mo + i];
    float r = ((double)P.mpViewCos[mo + i] > 0.998) ? 2.5f : 4.0f;               // :133-139
    if (P.th != 1.0f) r = __fmul_rn(r, P.th);
    const float rs = __fmul_rn(r, P.scaleFactors[min(max(lvl, 0), P.nlevels - 1)]);
    const float qx = P.mpX[mo + i], qy = P.mpY[mo + i], qxr = P.mpXR[mo + i];
    int c0, c1, r0, r1;
    if (!cell_range(P.g, qx, qy, rs, c0, c1, r0, r1)) return;
    const int minLevel = lvl - 1, maxLevel = lvl;
    const bool check = (minLevel > 0) || (maxLevel >= 0);
    for (int c = c0; c <= c1; c++) {
        const int s = cs[c * GRID_ROWS + r0], e = cs[c * GRID_ROWS + r1 + 1];
        for (int p = s; p < e; p++) {
            const uint4 rec = __ldg(cr + 3 * p);                                // {x, y, index, octave} in CSR order
            const int idx = (int)rec.z, o = (int)rec.w;
            if (check) {
                if (o < minLevel) continue;
                if (maxLevel >= 0 && o > maxLevel) continue;
            }
            if (!(fabsf(__fsub_rn(__uint_as_float(rec.x), qx)) < rs && fabsf(__fsub_rn(__uint_as_float(rec.y), qy)) < rs)) continue;
            if (ur && ur[idx] > 0) {                                            // :93-98
                const float er = fabsf(__fsub_rn(qxr, ur[idx]));
                if (er > rs) continue;
            }
            f(p, idx, o);
        }
    }
}

__global__ void __launch_bounds__(SP2_THREADS) k_search_proj2(const ProjParams P)
{
    constexpr int T = SP2_THREADS, U = 4;                      // U map points per thread and step: their list loads are in flight together
    const int tid = threadIdx.x;
    const int item = blockIdx.x;
    const int n = min(P.f.n[item], P.f.stride), nmp = min(P.mpN[item], P.mpStride);
    const int* ci = P.cellItems + (size_t)item * P.f.stride;
    const uint4* cr = P.cellRec + (size_t)item * P.f.stride * 3;
    int* kpmp = P.kpMp + (size_t)item * P.f.stride;
    const int* kpobs = P.kpMpObs ? P.kpMpObs + (size_t)item * P.f.stride : nullptr;
    const size_t mo = (size_t)item * P.mpStride;
    const uint4* md = reinterpret_cast<const uint4*>(P.mpDesc + mo * 32);
    const uint4* topk = P.topk + mo;
    const uint4* topkIdx = P.topkIdx + mo;
    const int* topkCount = P.topkCount + mo;

    extern __shared__ __align__(16) uint8_t sp2_smem[];
    const int nP = (P.f.stride + 15) & ~15, nL = (P.mpStride + 7) & ~7;
    // claim[k] = (0xffff - round) << 16 | lowest pending map point that lists keypoint k this round: a newer round's entries are
    // smaller than anything older, so the array never needs clearing
    unsigned* claim = reinterpret_cast<unsigned*>(sp2_smem);
    uint16_t* winner = reinterpret_cast<uint16_t*>(claim + nP);         // 1 + last accepting map point per keypoint (0: none)
    uint16_t* list[2] = {winner + nP, winner + nP + nL};                // pending map points, this round / next round
    uint8_t* occ = reinterpret_cast<uint8_t*>(list[1] + nL);            // keypoint holds a map point with observations (:89-91)
    __shared__ int s_len[2], s_nmatch;

    for (int k = tid; k < n; k += T) {
        const int held = kpmp[k];
        occ[k] = held != -1 && (held >= 0 ? P.mpObs[mo + held] : (kpobs ? kpobs[k] : 0)) > 0;
        winner[k] = 0;
        claim[k] = 0xffffffffu;
    }
    if (tid == 0) { s_len[0] = 0; s_len[1] = 0; s_nmatch = 0; }
    __syncthreads();
    for (int i = tid; i < nmp; i += T)
        if (topkCount[i] > 0) list[0][atomicAdd(&s_len[0], 1)] = (uint16_t)i;     // cnt <= 0: not in view / bad (:56-60) / no candidate (:73)
    __syncthreads();

    int cur = 0, acc = 0;
    unsigned stamp = 0xffffu << 16;
    while (true) {
        const int len = s_len[cur];
        if (len == 0) break;
        const uint16_t* L = list[cur];
        if (stamp == 0) {                                              // 65535 rounds used up (cannot happen with < 65536 map points; kept for safety)
            for (int k = tid; k < n; k += T) claim[k] = 0xffffffffu;
            stamp = 0xffffu << 16;
        }
        stamp -= 1u << 16;
        if (tid == 0) s_len[cur ^ 1] = 0;
        __syncthreads();
        // post: every pending map point on every unoccupied keypoint of its candidate set
        for (int b = tid; b < len; b += U * T) {
            int i[U], cnt[U];
            uint4 idv[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int pos = b + u * T;
                i[u] = pos < len ? (int)L[pos] : -1;
                cnt[u] = i[u] >= 0 ? topkCount[i[u]] : 0;
                idv[u] = i[u] >= 0 ? topkIdx[i[u]] : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (i[u] < 0) continue;
                if (cnt[u] <= 4) {
                    const int kid[4] = {(int)idv[u].x, (int)idv[u].y, (int)idv[u].z, (int)idv[u].w};
#pragma unroll
                    for (int j = 0; j < 4; j++) if (j < cnt[u] && !occ[kid[j]]) atomicMin(&claim[kid[j]], stamp | (unsigned)i[u]);
                } else {
                    const unsigned mine = stamp | (unsigned)i[u];
                    proj_for_each_candidate(P, item, mo, i[u], [&](int, int idx, int) { if (!occ[idx]) atomicMin(&claim[idx], mine); });
                }
            }
        }
        __syncthreads();
        // final map points decide and commit, the others go on next round's list
        for (int b = tid; b < len; b += U * T) {
            int i[U], cnt[U];
            uint4 idv[U], kk[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int pos = b + u * T;
                i[u] = pos < len ? (int)L[pos] : -1;
                cnt[u] = i[u] >= 0 ? topkCount[i[u]] : 0;
                idv[u] = i[u] >= 0 ? topkIdx[i[u]] : make_uint4(0, 0, 0, 0);
                kk[u] = i[u] >= 0 ? topk[i[u]] : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int ii = i[u], c = cnt[u];
                if (ii < 0) continue;
                const int kid[4] = {c > 0 ? (int)idv[u].x : -1, c > 1 ? (int)idv[u].y : -1, c > 2 ? (int)idv[u].z : -1, c > 3 ? (int)idv[u].w : -1};
                bool fin = true;
                const unsigned mine = stamp | (unsigned)ii;
                if (c <= 4) {
#pragma unroll
                    for (int j = 0; j < 4; j++) if (j < c && !occ[kid[j]] && claim[kid[j]] != mine) fin = false;
                } else {
                    proj_for_each_candidate(P, item, mo, ii, [&](int, int idx, int) { if (!occ[idx] && claim[idx] != mine) fin = false; });
                }
                if (!fin) { list[cur ^ 1][atomicAdd(&s_len[cur ^ 1], 1)] = (uint16_t)ii; continue; }
                const uint32_t key[4] = {kk[u].x, kk[u].y, kk[u].z, kk[u].w};
                Top2 t = {256, INT_MAX, -1, 256, INT_MAX, -1};
                int bestId = -1, found = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (found < 2 && j < c) {
                        const int dist = (int)(key[j] >> 23), o = (int)(key[j] & 31u);
                        if (!occ[kid[j]] && dist < 256) {
                            if (found == 0) { t.b = dist; t.ba = o; bestId = kid[j]; } else { t.s = dist; t.sa = o; }
                            found++;
                        }
                    }
                }
                const int d3 = (int)(key[3] >> 23);
                const bool resolved = found == 2 || c <= 4 || (found == 1 && t.b > TH_HIGH) || (found == 0 && d3 > TH_HIGH);
                if (!resolved) {        // the list is a truncation and too much of it is taken: rescan under the current occupancy
                    t = Top2{256, INT_MAX, -1, 256, INT_MAX, -1};
                    const uint4 a0 = __ldg(md + 2 * ii), a1 = __ldg(md + 2 * ii + 1);
                    proj_for_each_candidate(P, item, mo, ii, [&](int p, int idx, int o) {
                        if (occ[idx]) return;
                        top2_push(t, hamming256(a0, a1, __ldg(cr + 3 * p + 1), __ldg(cr + 3 * p + 2)), p, o);
                    });
                    bestId = t.bp >= 0 && t.b < 256 ? ci[t.bp] : -1;
                }
                if (t.b <= TH_HIGH && !(t.ba == t.sa && (float)t.b > __fmul_rn(P.nnratio, (float)t.s))) {   // :120-127
                    // :125, the last one in list order stays.  (Two map points that are final in the same round share no keypoint, and
                    // rounds are separated by barriers: a plain read-modify-write is race free)
                    if (winner[bestId] < ii + 1) winner[bestId] = (uint16_t)(ii + 1);
                    if (P.mpObs[mo + ii] > 0) occ[bestId] = 1;
                    acc++;
                }
            }
        }
        __syncthreads();
        cur ^= 1;
    }
    if (acc) atomicAdd(&s_nmatch, acc);
    __syncthreads();
    for (int k = tid; k < n; k += T) if (winner[k] != 0) kpmp[k] = (int)winner[k] - 1;
    if (tid == 0) P.nmatches[item] = s_nmatch;
}

int launch_build_grid(const FrameDev& f, const GridGeo& g, int* cellStart, int* cellItems, uint4* cellRec, int items, cudaStream_t st)
{
    k_build_grid<<<items, 256, 0, st>>>(f, g, cellStart, cellItems, cellRec);
    ORB_CHECK_LAUNCH("k_build_grid");
    return ORBB200_OK;
}

}  // namespace orbb200

// =========================================================================================
// host side
// =========================================================================================
using namespace orbb200;

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
    if (rc == ORBB200_OK) rc = m_alloc(m, (void**)&m->cellRec, sizeof(uint4) * 3 * np);
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
extern "C" int orbb200_matcher_device(const orbb200_matcher* m) { return m ? m->device : ORBB200_EINVAL; }
extern "C" int orbb200_matcher_sync(orbb200_matcher* m)
{
    if (!m) { set_error("null handle"); return ORBB200_EINVAL; }
    ORB_CUDA(cudaSetDevice(m->device));
    ORB_CUDA(cudaStreamSynchronize(m->stream));
    return ORBB200_OK;
}

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
    // smallest second-best distance that passes `best < nnratio * second` for every best <= TH_LOW, in the float arithmetic of :465
    P.farDist = 257;
    for (int d = TH_LOW + 1; d <= 256; d++)
        if ((float)TH_LOW < (float)d * nnratio) { P.farDist = d; break; }
    // scratch strides follow the views
    k_build_grid<<<items, 256, 0, st>>>(P.f2, P.g, m->cellStart, m->cellItems, nullptr);
    ORB_CHECK_LAUNCH("k_build_grid");
    k_init_topk<<<dim3((f1->stride + orbb200::TOPK_QPB - 1) / orbb200::TOPK_QPB, items), 128, 0, st>>>(P);
    ORB_CHECK_LAUNCH("k_init_topk");
    {
        const size_t sm = 4 * 3 * (size_t)((f2->stride + 15) & ~15);
        if (sm > 200 * 1024) { set_error("more than %d keypoints per frame", 16 * 1024); return ORBB200_EINVAL; }
        ORB_CUDA(ensure_dynamic_smem((const void*)k_search_init, m->device, sm));
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
    P.mpStride = mp->stride; P.g = grid_geo(bounds); P.cellStart = m->cellStart; P.cellItems = m->cellItems; P.cellRec = m->cellRec;
    P.nlevels = nlevels; P.nmatches = dN; P.items = items; P.nnratio = nnratio; P.th = th;
    P.topk = m->topk; P.topkCount = m->topkCount; P.topkIdx = m->topkIdx;
    k_build_grid<<<items, 256, 0, st>>>(P.f, P.g, m->cellStart, m->cellItems, m->cellRec);
    ORB_CHECK_LAUNCH("k_build_grid");
    k_proj_topk<<<dim3((mp->stride + 127) / 128, items), 128, 0, st>>>(P);
    ORB_CHECK_LAUNCH("k_proj_topk");
    {
        // frame-wide reservations when the per-frame state fits a CTA's shared memory comfortably, else one warp per frame
        const size_t nP = (size_t)((f->stride + 15) & ~15), sm2 = 7 * nP + 4 * (size_t)((mp->stride + 7) & ~7) + 16;
        static const bool oldResolve = getenv("ORBB200_PROJ_RESOLVE") && atoi(getenv("ORBB200_PROJ_RESOLVE")) == 1;
        if (sm2 <= 100 * 1024 && mp->stride < 65535 && !oldResolve) {
            ORB_CUDA(ensure_dynamic_smem((const void*)k_search_proj2, m->device, sm2));
            k_search_proj2<<<items, SP2_THREADS, sm2, st>>>(P);
        } else {
            const size_t sm = 8 * (size_t)((f->stride + 15) & ~15);
            if (sm > 200 * 1024) { set_error("more than %d keypoints per frame", 25 * 1024); return ORBB200_EINVAL; }
            ORB_CUDA(ensure_dynamic_smem((const void*)k_search_proj, m->device, sm));
            k_search_proj<<<(items + 3) / 4, 128, sm, st>>>(P);
        }
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
