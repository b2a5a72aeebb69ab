// orb_matcher_proj.cu -- the ORBmatcher searches whose projection runs on the device:
//   SearchByProjection(CurrentFrame, LastFrame, th, bMono)              S/ORBmatcher.cc:1332-1474
//   SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist)    :1476-1603
//   SearchByProjection(pKF, Scw, vpPoints, vpMatched, th)               :294-407
//   the candidate search of Fuse(pKF, vpMapPoints, th) / Fuse(pKF, Scw, ...) and the legs of SearchBySim3   :829-1330
// (S/ = oRB_SLAM2_Android/src/main/jni/ORB_SLAM2/src/).  Same two-phase scheme as orb_matcher.cu for the greedy ones.
#include <math_constants.h>
#include "matcher_common.cuh"
#include "../../include/orb_b200_logf.inc"

namespace orbb200 {

// =========================================================================================
// SearchByProjection(CurrentFrame, LastFrame, th, bMono)  (S/ORBmatcher.cc:1332-1474), SURVEY 8(f) N2
// =========================================================================================
struct LastParams {
    FrameDev f;                 // current frame
    const float* uRight;        // items x f.stride or NULL
    GridGeo g;
    const int* cellStart;
    const int* cellItems;
    const uint4* cellRec;       // the grid's keypoints in CSR order: {x, y, index, octave} + descriptor (k_build_grid)
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
__device__ __forceinline__ bool last_candidate(const LastParams& P, const LastQuery& q, int idx, float x, float y, int o, const float* ur)
{
    if ((q.minLevel > 0) || (q.maxLevel >= 0)) {
        if (o < q.minLevel) return false;
        if (q.maxLevel >= 0 && o > q.maxLevel) return false;
    }
    if (!(fabsf(__fsub_rn(x, q.u)) < q.radius && fabsf(__fsub_rn(y, q.v)) < q.radius)) return false;
    if (ur && ur[idx] > 0) {
        const float pr = __fsub_rn(q.u, __fmul_rn(P.mbf, q.invzc));
        if (fabsf(__fsub_rn(pr, ur[idx])) > q.radius) return false;
    }
    return true;
}
__device__ __forceinline__ bool last_candidate(const LastParams& P, const LastQuery& q, int idx, const float* kx, const float* ky,
                                               const int* koct, const float* ur)
{
    return last_candidate(P, q, idx, kx[idx], ky[idx], koct[idx], ur);
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
        const float* ur = P.uRight ? P.uRight + (size_t)item * P.f.stride : nullptr;
        const int* cs = P.cellStart + (size_t)item * (GRID_CELLS + 1);
        const int* kpmp = P.kpMp + (size_t)item * P.f.stride;
        const int* kpobs = P.kpMpObs ? P.kpMpObs + (size_t)item * P.f.stride : nullptr;
        const uint4* md = reinterpret_cast<const uint4*>(P.mpDesc + lo * 32);
        const uint4 a0 = __ldg(md), a1 = __ldg(md + 1);
        const uint4* cr = P.cellRec + (size_t)item * P.f.stride * 3;
        for (int c = c0; c <= c1; c++) {
            const int s = cs[c * GRID_ROWS + r0], e = cs[c * GRID_ROWS + r1 + 1];
            for (int p = s; p < e; p++) {
                const uint4 rec = __ldg(cr + 3 * p);                             // {x, y, index, octave} in CSR order
                const uint4 b0 = __ldg(cr + 3 * p + 1), b1 = __ldg(cr + 3 * p + 2);
                const int idx = (int)rec.z;
                if (!last_candidate(P, q, idx, __uint_as_float(rec.x), __uint_as_float(rec.y), (int)rec.w, ur)) continue;
                const int held = kpmp[idx];                                      // initial occupancy (:1409-1411, :1546-1547)
                if (held != -1 && (P.kind != 0 || (held >= 0 ? P.mpObs[(size_t)item * P.lastStride + held] : (kpobs ? kpobs[idx] : 0)) > 0)) continue;
                const int dist = hamming256(a0, a1, b0, b1);
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
    const float* kang = P.f.angle + (size_t)item * P.f.stride;
    const float* ur = P.uRight ? P.uRight + (size_t)item * P.f.stride : nullptr;
    const int* cs = P.cellStart + (size_t)item * (GRID_CELLS + 1);
    const int* ci = P.cellItems + (size_t)item * P.f.stride;
    const uint4* cr = P.cellRec + (size_t)item * P.f.stride * 3;
    int* kpmp = P.kpMp + (size_t)item * P.f.stride;
    const int* kpobs = P.kpMpObs ? P.kpMpObs + (size_t)item * P.f.stride : nullptr;
    const size_t lo = (size_t)item * P.lastStride;
    int* hbin = P.histBin + lo;
    int* hidx = P.histIdx + lo;

    extern __shared__ uint8_t s_occ_all[];
    uint8_t* occ = s_occ_all + (size_t)(threadIdx.x >> 5) * 2 * ((P.f.stride + 15) & ~15);
    uint8_t* claim = occ + ((P.f.stride + 15) & ~15);
    for (int idx = lane; idx < n; idx += 32) {
        const int held = kpmp[idx];
        occ[idx] = held != -1 && (P.kind != 0 || (held >= 0 ? P.mpObs[lo + held] : (kpobs ? kpobs[idx] : 0)) > 0);
        claim[idx] = 0xff;
    }
    for (int i = lane; i < nl; i += 32) hbin[i] = -1;
    __syncwarp();

    int nmatches = 0;
    const int accept = P.kind == 1 ? P.orbDist : (P.kind == 2 ? TH_LOW : TH_HIGH);
    // Lane j owns last-frame point base + j and decides it SPECULATIVELY against the current occupancy, all 32 at
    // once; a decision is final when no earlier point of the batch takes a keypoint the lane looked at (its <= 4 list
    // entries).  The longest clean prefix is committed in parallel, the rest is decided again.
    for (int base = 0; base < nl; base += 32) {
      const int mine = base + lane;
      int cnt = -1, obs = 0;
      float ang = 0.f;
      uint4 kk = make_uint4(0, 0, 0, 0), idv = kk;
      if (mine < nl) { cnt = P.topkCount[lo + mine]; kk = P.topk[lo + mine]; idv = P.topkIdx[lo + mine]; obs = P.kind != 0 ? 1 : P.mpObs[lo + mine]; ang = P.lastAng[lo + mine]; }
      unsigned todo = __ballot_sync(0xffffffffu, cnt > 0);
      const uint32_t key[4] = {kk.x, kk.y, kk.z, kk.w};
      const int kid[4] = {cnt > 0 ? (int)idv.x : -1, cnt > 1 ? (int)idv.y : -1, cnt > 2 ? (int)idv.z : -1, cnt > 3 ? (int)idv.w : -1};
      while (todo) {
        const int first = __ffs(todo) - 1;
        const bool pending = (todo >> lane) & 1u;
        // only the best candidate matters here (no ratio test): the first entry of the list whose keypoint is free
        int bestDist = 256, bestIdx = -1;
        bool resolved = true;
        if (pending) {
            bool found = false;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                if (!found && e < cnt) {
                    const int dist = (int)(key[e] >> 23);
                    if (!occ[kid[e]] && dist < 256) { bestDist = dist; bestIdx = kid[e]; found = true; }
                }
            }
            resolved = found || cnt <= 4 || (int)(key[3] >> 23) > accept;
        }
        bool take;
        int stop;
        if (__shfl_sync(0xffffffffu, (int)!resolved, first)) {
            // the lowest pending point sees the exact state and every listed keypoint was taken: rescan all candidates
            const int i = base + first;
            const LastQuery q = last_query(P, item, i);
            int c0, c1, r0, r1;
            cell_range(P.kind == 2 ? P.q : P.g, q.u, q.v, q.radius, c0, c1, r0, r1);
            const uint4* md = reinterpret_cast<const uint4*>(P.mpDesc + (lo + i) * 32);
            const uint4 a0 = __ldg(md), a1 = __ldg(md + 1);
            int bd = 256, bp = INT_MAX;
            for (int c = c0; c <= c1; c++) {
                const int s = cs[c * GRID_ROWS + r0], e = cs[c * GRID_ROWS + r1 + 1];
                for (int p = s + lane; p < e; p += 32) {
                    const uint4 rec = __ldg(cr + 3 * p);                         // {x, y, index, octave} in CSR order
                    const uint4 b0 = __ldg(cr + 3 * p + 1), b1 = __ldg(cr + 3 * p + 2);
                    const int idx = (int)rec.z;
                    if (!last_candidate(P, q, idx, __uint_as_float(rec.x), __uint_as_float(rec.y), (int)rec.w, ur) || occ[idx]) continue;
                    const int dist = hamming256(a0, a1, b0, b1);
                    if (dist < bd) { bd = dist; bp = p; }
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const int od = __shfl_xor_sync(0xffffffffu, bd, d), op = __shfl_xor_sync(0xffffffffu, bp, d);
                if (key_lt(od, op, bd, bp)) { bd = od; bp = op; }
            }
            if (lane == first) { bestDist = bd; bestIdx = bd < 256 ? ci[bp] : -1; }
            take = lane == first && bd <= accept;
            stop = first + 1;
        } else {
            take = pending && resolved && bestDist <= accept;                     // :1436-1452, :1561-1579
            const int w = (take && obs > 0) ? bestIdx : -1;        // the occupancy this point would set
            stop = clean_prefix(claim, lane, w, kid, pending && !resolved);
        }
        const bool commit = take && lane < stop;
        nmatches += __popc(__ballot_sync(0xffffffffu, commit));
        // several points without observations may take the same keypoint; the last one in list order stays
        const unsigned same = __match_any_sync(0xffffffffu, commit ? bestIdx : -1 - lane);
        if (commit) {
            if (lane == 31 - __clz(same)) kpmp[bestIdx] = mine;
            if (obs > 0) occ[bestIdx] = 1;
            if (P.checkOri) {
                float rot = __fsub_rn(ang, kang[bestIdx]);
                if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));
                if (bin == HISTO_LENGTH) bin = 0;
                hbin[mine] = bin; hidx[mine] = bestIdx;
            }
        }
        todo = stop < 32 ? (todo & (0xffffffffu << stop)) : 0u;
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

}  // namespace orbb200

// =========================================================================================
// host side
// =========================================================================================
using namespace orbb200;

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
    P.lastStride = last->stride; P.g = grid_geo(bounds); P.cellStart = m->cellStart; P.cellItems = m->cellItems; P.cellRec = m->cellRec;
    P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3]; P.mbf = mbf;
    P.minX = bounds[0]; P.minY = bounds[1]; P.maxX = bounds[2]; P.maxY = bounds[3];
    P.nmatches = dN; P.items = items; P.mode = mode; P.checkOri = check_orientation; P.th = th;
    P.topk = m->topk; P.topkCount = m->topkCount; P.topkIdx = m->topkIdx; P.histBin = m->scratchA; P.histIdx = m->scratchB;
    if ((rc = launch_build_grid(P.f, P.g, m->cellStart, m->cellItems, m->cellRec, items, st))) return rc;
    k_last_topk<<<dim3((last->stride + 127) / 128, items), 128, 0, st>>>(P);
    ORB_CHECK_LAUNCH("k_last_topk");
    {
        const size_t sm = 8 * (size_t)((cur->stride + 15) & ~15);
        if (sm > 200 * 1024) { set_error("more than %d keypoints per frame", 25 * 1024); return ORBB200_EINVAL; }
        ORB_CUDA(ensure_dynamic_smem((const void*)k_search_last, m->device, sm));
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
    P.lastStride = kf->stride; P.g = grid_geo(bounds); P.cellStart = m->cellStart; P.cellItems = m->cellItems; P.cellRec = m->cellRec;
    P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3];
    P.minX = bounds[0]; P.minY = bounds[1]; P.maxX = bounds[2]; P.maxY = bounds[3];
    P.nmatches = dN; P.items = items; P.mode = 0; P.checkOri = check_orientation; P.th = th;
    P.topk = m->topk; P.topkCount = m->topkCount; P.topkIdx = m->topkIdx; P.histBin = m->scratchA; P.histIdx = m->scratchB;
    if ((rc = launch_build_grid(P.f, P.g, m->cellStart, m->cellItems, m->cellRec, items, st))) return rc;
    k_last_topk<<<dim3((kf->stride + 127) / 128, items), 128, 0, st>>>(P);
    ORB_CHECK_LAUNCH("k_last_topk");
    {
        const size_t sm = 8 * (size_t)((cur->stride + 15) & ~15);
        if (sm > 200 * 1024) { set_error("more than %d keypoints per frame", 25 * 1024); return ORBB200_EINVAL; }
        ORB_CUDA(ensure_dynamic_smem((const void*)k_search_last, m->device, sm));
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
    if ((rc = launch_build_grid(P.f, P.g, m->cellStart, m->cellItems, m->cellRec, items, st))) return rc;
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
    P.cellStart = m->cellStart; P.cellItems = m->cellItems; P.cellRec = m->cellRec;
    P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3];
    P.nmatches = dN; P.items = items; P.mode = 0; P.checkOri = 0; P.th = (float)th;
    P.topk = m->topk; P.topkCount = m->topkCount; P.topkIdx = m->topkIdx; P.histBin = m->scratchA; P.histIdx = m->scratchB;
    if ((rc = launch_build_grid(P.f, P.g, m->cellStart, m->cellItems, m->cellRec, items, st))) return rc;
    k_last_topk<<<dim3((pts->stride + 127) / 128, items), 128, 0, st>>>(P);
    ORB_CHECK_LAUNCH("k_last_topk");
    {
        const size_t sm = 8 * (size_t)((kf->stride + 15) & ~15);
        if (sm > 200 * 1024) { set_error("more than %d keypoints per frame", 25 * 1024); return ORBB200_EINVAL; }
        ORB_CUDA(ensure_dynamic_smem((const void*)k_search_last, m->device, sm));
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
