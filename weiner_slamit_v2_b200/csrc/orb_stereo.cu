// orb_stereo.cu -- Frame::ComputeStereoMatches (S/Frame.cc:591-763) for batches of rectified stereo pairs on B200
// (the one consumer of ORBextractor::mvImagePyramid outside the extractor).
//
// Reference, per left keypoint: the right keypoints whose row band [floor(y - r), ceil(y + r)], r = 2 * scale[octave],
// contains the left keypoint's row (vRowIndices, :598-617), within one octave and the disparity range, closest
// descriptor first-wins (:631-672); then an 11 x 11 sum of absolute differences of centre-subtracted patches slid over
// +-5 columns of the right pyramid level (:680-713), a parabola through the three costs around the minimum (:720-729),
// depth = mbf / disparity (:731-745); finally every match whose patch cost reaches 1.5 * 1.4 * median is withdrawn
// (:749-762).  Left keypoints are independent of each other up to that last step.
//
// Device mapping: k_stereo_match, one WARP per left keypoint.  The right keypoints' row bands, octaves and columns sit
// in shared memory (8 bytes per keypoint); the lanes stride over them, which visits them in index order = the
// push_back order of every row list, so "first smallest distance" is one packed (distance, index) warp minimum and no
// row table is needed.  The patch cost uses integers (the reference's floats hold small integers, so cv::norm's double
// sum is exact): lane p handles patch pixels p, p + 32, ... for all 11 shifts, eleven warp sums.  k_stereo_median, one
// CTA per pair: the median cost by a two-pass byte histogram (costs are < 2^16), then the withdrawal.
//
// Where the reference would leave an image (cv::Mat::rowRange / colRange throw, vRowIndices indexed out of range) the
// keypoint simply gets no depth; keypoints produced by the extractor (>= 19 px inside their level) never get there.
#include "matcher_common.cuh"

namespace orbb200 {

constexpr int STEREO_MAXL = ORBB200_MAX_LEVELS;

struct PyrDev {
    const uint8_t* level[STEREO_MAXL];
    size_t frameStride[STEREO_MAXL];
    int pitch[STEREO_MAXL];
};

struct StereoParams {
    FrameDev l, r;
    PyrDev lp, rp;
    int w[STEREO_MAXL], h[STEREO_MAXL];
    float scale[STEREO_MAXL], invScale[STEREO_MAXL];
    int nlevels;
    float mb, mbf;
    float* uRight;          // items x l.stride
    float* depth;           // items x l.stride
    int* cost;              // items x l.stride: patch cost of a match, -1 = none
    int* nmatches;          // items
};

__global__ void __launch_bounds__(256) k_stereo_match(const StereoParams P)
{
    extern __shared__ uint32_t s_right[];                 // [stride] minr | maxr << 13 | octave << 26, then [stride] uR
    const int item = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = min(P.l.n[item], P.l.stride), nr = min(P.r.n[item], P.r.stride);
    uint32_t* s_band = s_right;
    float* s_u = reinterpret_cast<float*>(s_right + P.r.stride);
    const size_t ro = (size_t)item * P.r.stride, lo = (size_t)item * P.l.stride;
    const int nRows = P.h[0];                                                      // :596
    for (int i = tid; i < nr; i += 256) {
        const float y = P.r.y[ro + i];
        const int oct = P.r.octave[ro + i];
        const float r = __fmul_rn(2.0f, P.scale[min(max(oct, 0), P.nlevels - 1)]);   // :608
        int maxr = (int)ceilf(__fadd_rn(y, r)), minr = (int)floorf(__fsub_rn(y, r));
        minr = max(minr, 0); maxr = min(maxr, nRows - 1);
        if (maxr < minr || oct < 0 || oct >= P.nlevels) { minr = 1; maxr = 0; }
        s_band[i] = (uint32_t)minr | ((uint32_t)maxr << 13) | ((uint32_t)(oct & 31) << 26);
        s_u[i] = P.r.x[ro + i];
    }
    __syncthreads();

    const float minZ = P.mb, maxD = __fdiv_rn(P.mbf, minZ);                        // :620-622
    const uint4* dl = reinterpret_cast<const uint4*>(P.l.desc + lo * 32);
    const uint4* dr = reinterpret_cast<const uint4*>(P.r.desc + ro * 32);
    const int warpsPerItem = gridDim.x * 8;
    for (int iL = blockIdx.x * 8 + warp; iL < n; iL += warpsPerItem) {
        float outU = -1.0f, outD = -1.0f;
        int outCost = -1;
        const int levelL = P.l.octave[lo + iL];
        const float vL = P.l.y[lo + iL], uL = P.l.x[lo + iL];
        const float minU = __fsub_rn(uL, maxD), maxU = __fsub_rn(uL, -3.0f);
        uint32_t bestKey = 0xffffffffu;
        if (vL >= 0.0f && vL < (float)nRows && !(maxU < 0.0f) && levelL >= 0 && levelL < P.nlevels) {   // :634-644
            const int row = (int)vL;
            const uint4 a0 = __ldg(dl + 2 * iL), a1 = __ldg(dl + 2 * iL + 1);
            for (int iR = lane; iR < nr; iR += 32) {
                const uint32_t b = s_band[iR];
                const int minr = (int)(b & 0x1fffu), maxr = (int)((b >> 13) & 0x1fffu), oR = (int)(b >> 26);
                if (row < minr || row > maxr) continue;
                if (oR < levelL - 1 || oR > levelL + 1) continue;                 // :656-657
                const float uR = s_u[iR];
                if (!(uR >= minU && uR <= maxU)) continue;                        // :661
                const int dist = hamming256(a0, a1, __ldg(dr + 2 * iR), __ldg(dr + 2 * iR + 1));
                if (dist < TH_HIGH) bestKey = min(bestKey, ((uint32_t)dist << 16) | (uint32_t)iR);   // first smallest (:666-670)
            }
        }
        bestKey = __reduce_min_sync(0xffffffffu, bestKey);
        if (bestKey != 0xffffffffu) {                                             // :675
            const int bestIdxR = (int)(bestKey & 0xffffu);
            const float uR0 = s_u[bestIdxR];
            const float sf = P.invScale[levelL];
            const float scaleduL = roundf(__fmul_rn(uL, sf)), scaledvL = roundf(__fmul_rn(vL, sf));
            const float scaleduR0 = roundf(__fmul_rn(uR0, sf));
            const int W = P.w[levelL], H = P.h[levelL];
            const bool inside = scaledvL - 5.0f >= 0.0f && scaledvL + 6.0f <= (float)H && scaleduL - 5.0f >= 0.0f &&
                                scaleduL + 6.0f <= (float)W && scaleduR0 - 10.0f >= 0.0f;
            const float iniu = scaleduR0, endu = __fadd_rn(scaleduR0, 11.0f);      // scaleduR0 + L - w, + L + w + 1 (:693-694)
            if (inside && !(iniu < 0.0f || endu >= (float)W)) {                    // :695-696
                const int v0 = (int)scaledvL, u0 = (int)scaleduL, r0 = (int)scaleduR0;
                const uint8_t* Lp = P.lp.level[levelL] + (size_t)item * P.lp.frameStride[levelL];
                const uint8_t* Rp = P.rp.level[levelL] + (size_t)item * P.rp.frameStride[levelL];
                const int lpitch = P.lp.pitch[levelL], rpitch = P.rp.pitch[levelL];
                const int cL = Lp[(size_t)v0 * lpitch + u0];
                // centre of the right patch at shift s = lane - 5 (lanes 0..10)
                const int cRmine = lane < 11 ? (int)Rp[(size_t)v0 * rpitch + r0 + lane - 5] : 0;
                int sad[11], cR[11];
#pragma unroll
                for (int s = 0; s < 11; s++) { sad[s] = 0; cR[s] = __shfl_sync(0xffffffffu, cRmine, s); }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int p = lane + 32 * k;
                    if (p < 121) {
                        const int dy = p / 11 - 5, dx = p - (p / 11) * 11 - 5;
                        const int a = (int)Lp[(size_t)(v0 + dy) * lpitch + u0 + dx] - cL;
                        const uint8_t* rrow = Rp + (size_t)(v0 + dy) * rpitch + r0 + dx - 5;
#pragma unroll
                        for (int s = 0; s < 11; s++) {
                            const int b = (int)rrow[s] - cR[s];
                            sad[s] += abs(a - b);
                        }
                    }
                }
                int best = INT_MAX, bestinc = 0;
                int d[11];
#pragma unroll
                for (int s = 0; s < 11; s++) {
                    d[s] = __reduce_add_sync(0xffffffffu, sad[s]);
                    if (d[s] < best) { best = d[s]; bestinc = s - 5; }             // :706-710
                }
                if (bestinc != -5 && bestinc != 5) {                              // :715-716
                    float dist1 = 0.f, dist2 = 0.f, dist3 = 0.f;
#pragma unroll
                    for (int s = 1; s < 10; s++)
                        if (s - 5 == bestinc) { dist1 = (float)d[s - 1]; dist2 = (float)d[s]; dist3 = (float)d[s + 1]; }
                    const float deltaR = __fdiv_rn(__fsub_rn(dist1, dist3),
                                                   __fmul_rn(2.0f, __fsub_rn(__fadd_rn(dist1, dist3), __fmul_rn(2.0f, dist2))));
                    if (!(deltaR < -1.0f || deltaR > 1.0f)) {                     // :725-726
                        float bestuR = __fmul_rn(P.scale[levelL], __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));
                        float disparity = __fsub_rn(uL, bestuR);
                        if (disparity >= 0.0f && disparity < maxD) {              // :733-744
                            if (disparity <= 0.0f) { disparity = (float)0.01; bestuR = (float)__dsub_rn((double)uL, 0.01); }
                            outD = __fdiv_rn(P.mbf, disparity);
                            outU = bestuR;
                            outCost = best;
                        }
                    }
                }
            }
        }
        if (lane == 0) { P.uRight[lo + iL] = outU; P.depth[lo + iL] = outD; P.cost[lo + iL] = outCost; }
    }
}

// Withdraws the matches whose patch cost reaches 1.5f * 1.4f * median (:749-762); one CTA per pair.
__global__ void __launch_bounds__(256) k_stereo_median(const StereoParams P)
{
    __shared__ int hist[256];
    __shared__ int s_sel[3];                  // number of matches, selected high byte, rank inside it
    const int item = blockIdx.x, tid = threadIdx.x;
    const int n = min(P.l.n[item], P.l.stride);
    const size_t lo = (size_t)item * P.l.stride;
    const int* cost = P.cost + lo;
    hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 256) { const int c = cost[i]; if (c >= 0) atomicAdd(&hist[(c >> 8) & 255], 1); }
    __syncthreads();
    if (tid == 0) {
        int nd = 0;
        for (int b = 0; b < 256; b++) nd += hist[b];
        int k = nd / 2, hb = 0;                // vDistIdx[vDistIdx.size() / 2] (:750)
        while (hb < 255 && k >= hist[hb]) { k -= hist[hb]; hb++; }
        s_sel[0] = nd; s_sel[1] = hb; s_sel[2] = k;
    }
    __syncthreads();
    const int nd = s_sel[0], hb = s_sel[1], k = s_sel[2];
    if (nd == 0) { if (tid == 0) P.nmatches[item] = 0; return; }
    __syncthreads();
    hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 256) { const int c = cost[i]; if (c >= 0 && ((c >> 8) & 255) == hb) atomicAdd(&hist[c & 255], 1); }
    __syncthreads();
    if (tid == 0) {
        int kk = k, lb = 0;
        while (lb < 255 && kk >= hist[lb]) { kk -= hist[lb]; lb++; }
        s_sel[1] = (hb << 8) | lb;
    }
    __syncthreads();
    const float median = (float)s_sel[1];
    const float thDist = __fmul_rn(1.5f * 1.4f, median);                           // :751
    int kept = 0;
    for (int i = tid; i < n; i += 256) {
        const int c = cost[i];
        if (c < 0) continue;
        if ((float)c < thDist) kept++;
        else { P.uRight[lo + i] = -1.0f; P.depth[lo + i] = -1.0f; }               // :753-761
    }
    __syncthreads();
    if (tid == 0) s_sel[0] = 0;
    __syncthreads();
    if (kept) atomicAdd(&s_sel[0], kept);
    __syncthreads();
    if (tid == 0) P.nmatches[item] = s_sel[0];
}

}  // namespace orbb200

using namespace orbb200;

static int pyr_to_dev(Stager& s, const orbb200_pyramid_view* v, int items, int nlevels, bool onDevice, PyrDev* d, const int* w, const int* h)
{
    for (int l = 0; l < nlevels; l++) {
        if (!v->level[l] || v->width[l] != w[l] || v->height[l] != h[l] || v->pitch[l] < v->width[l]) {
            set_error("pyramid level %d: missing or of a different size than the left level", l);
            return ORBB200_EINVAL;
        }
        if (onDevice) { d->level[l] = v->level[l]; d->frameStride[l] = v->frame_stride[l]; d->pitch[l] = v->pitch[l]; continue; }
        const size_t pitch = align_up((size_t)w[l], 16), frame = pitch * (size_t)h[l];
        uint8_t* dst = s.out<uint8_t>(frame * items);
        for (int i = 0; i < items; i++)
            ORB_CUDA(cudaMemcpy2DAsync(dst + frame * i, pitch, v->level[l] + v->frame_stride[l] * i, (size_t)v->pitch[l],
                                       (size_t)w[l], (size_t)h[l], cudaMemcpyHostToDevice, s.st));
        d->level[l] = dst; d->frameStride[l] = frame; d->pitch[l] = (int)pitch;
    }
    return ORBB200_OK;
}

extern "C" int orbb200_compute_stereo_matches(orbb200_matcher* m, int items, const orbb200_frame_view* left,
                                              const orbb200_frame_view* right, const orbb200_pyramid_view* lpyr,
                                              const orbb200_pyramid_view* rpyr, const float* scale_factors,
                                              const float* inv_scale_factors, int nlevels, float mb, float mbf,
                                              float* u_right, float* depth, int32_t* nmatches, int on_device)
{
    if (!m || !left || !right || !lpyr || !rpyr || !scale_factors || !inv_scale_factors || !u_right || !depth || !nmatches) {
        set_error("null argument"); return ORBB200_EINVAL;
    }
    if (!left->n || !left->x || !left->y || !left->octave || !left->desc || !right->n || !right->x || !right->y ||
        !right->octave || !right->desc) { set_error("incomplete view"); return ORBB200_EINVAL; }
    int rc;
    if ((rc = check_view(m, items, left->stride, "left frame")) || (rc = check_view(m, items, right->stride, "right frame"))) return rc;
    if (nlevels < 1 || nlevels > STEREO_MAXL || lpyr->nlevels < nlevels || rpyr->nlevels < nlevels) { set_error("bad number of pyramid levels"); return ORBB200_EINVAL; }
    if (right->stride > 65535) { set_error("more than 65535 right keypoints per frame"); return ORBB200_EINVAL; }
    if (lpyr->height[0] > 8191) { set_error("images taller than 8191 rows"); return ORBB200_EINVAL; }
    // (mb <= 0 is not an error: the reference's stereo constructor calls ComputeStereoMatches BEFORE it assigns mb,
    // S/Frame.cc:104 against :130, so the first frame runs with whatever the storage held; mb = 0 gives maxD = +inf in
    // IEEE arithmetic, here as there)
    const bool devViews = (on_device & ORBB200_DEVICE_VIEWS) != 0, devPyr = (on_device & ORBB200_DEVICE_PYRAMIDS) != 0;
    ORB_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = m->stream;
    StereoParams P;
    memset(&P, 0, sizeof(P));
    for (int l = 0; l < nlevels; l++) {
        P.w[l] = lpyr->width[l]; P.h[l] = lpyr->height[l];
        P.scale[l] = scale_factors[l]; P.invScale[l] = inv_scale_factors[l];
        if (P.w[l] < 1 || P.h[l] < 1) { set_error("empty pyramid level %d", l); return ORBB200_EINVAL; }
    }
    P.nlevels = nlevels; P.mb = mb; P.mbf = mbf;
    const size_t npl = (size_t)items * left->stride;
    Stager s{m, 0, st};
    size_t bytes = pad(npl * 4);                                                   // cost
    if (!devViews) bytes += frame_bytes(left, items) + frame_bytes(right, items) + 2 * pad(npl * 4) + pad((size_t)items * 4);
    if (!devPyr)
        for (int l = 0; l < nlevels; l++) bytes += 2 * pad(align_up((size_t)P.w[l], 16) * P.h[l] * items);
    if ((rc = s.reserve(bytes))) return rc;
    P.cost = s.out<int>(npl);
    int* dN;
    if (devViews) {
        P.l = as_dev(left); P.r = as_dev(right); P.uRight = u_right; P.depth = depth; dN = nmatches;
    } else {
        if ((rc = upload_frame(s, left, items, &P.l, false)) || (rc = upload_frame(s, right, items, &P.r, false))) return rc;
        P.uRight = s.out<float>(npl); P.depth = s.out<float>(npl); dN = s.out<int>(items);
    }
    P.nmatches = dN;
    if ((rc = pyr_to_dev(s, lpyr, items, nlevels, devPyr, &P.lp, P.w, P.h)) || (rc = pyr_to_dev(s, rpyr, items, nlevels, devPyr, &P.rp, P.w, P.h))) return rc;
    const size_t sm = 8 * (size_t)right->stride;
    if (sm > 200 * 1024) { set_error("more than %d right keypoints per frame", 25 * 1024); return ORBB200_EINVAL; }
    ORB_CUDA(ensure_dynamic_smem((const void*)k_stereo_match, m->device, sm));
    const int chunks = std::max(1, std::min((left->stride + 63) / 64, (148 * 8 + items - 1) / items));
    k_stereo_match<<<dim3(chunks, items), 256, sm, st>>>(P);
    ORB_CHECK_LAUNCH("k_stereo_match");
    k_stereo_median<<<items, 256, 0, st>>>(P);
    ORB_CHECK_LAUNCH("k_stereo_median");
    m->lastLaunches = 2;
    if (!devViews) {
        ORB_CUDA(cudaMemcpyAsync(u_right, P.uRight, npl * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaMemcpyAsync(depth, P.depth, npl * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaMemcpyAsync(nmatches, dN, (size_t)items * 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaStreamSynchronize(st));
    }
    return ORBB200_OK;
}
