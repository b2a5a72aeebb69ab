// orb_extractor.cu -- ORBextractor::operator() for batches of frames on B200 (sm_100a).
//
// Reference path: S/ORBextractor.cc (S/ = oRB_SLAM2_Android/src/main/jni/ORB_SLAM2/src/):
//   ComputePyramid :1138-1168, ComputeKeyPointsOctTree :778-873, DistributeOctTree :552-776,
//   IC_Angle :82-109, GaussianBlur call :1117, computeOrbDescriptor :113-152, operator() :1064-1136.
// Everything here is integer / byte streaming work (HBM- and issue-bound): no tensor cores.
//
// Kernel map (one launch each per batch unless noted; round 1's k_resize / k_fast / k_describe stay selectable for A/B runs):
//   k_resize3    x (nlevels-1)  pyramid level l from level l-1, 11-bit fixed-point bilinear; a warp = 64 columns x 16 rows of two
//                frames behind its own TMA window, one pass over the source rows; the launches are chained (griddepcontrol)
//   k_fast2      one warp = one 30-px detection cell (TMA window, 16-bit tile): FAST-9/16 in three passes, 3x3 NMS, iniThFAST ->
//                minThFAST retry decided per cell, candidates appended per (frame,level); two launches when some cells are wide
//   k_quadtree   one CTA per (frame,level): DistributeOctTree as parallel rounds over a node table
//   k_blur       7x7 fixed-point separable Gaussian, persistent warps over 128 x 24 tiles (one TMA window per tile)
//   k_describe3  persistent warps over groups of 8 keypoints: IC_Angle moments + steered BRIEF + output, TMA patches
#include <cuda.h>          // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <utility>
#include <vector>
#include "common.cuh"

namespace orbb200 {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

constexpr int MAXL = 16;             // pyramid levels supported per handle
constexpr int GRAPH_MAX_BATCH = 8;   // host-path calls up to this many frames replay a captured CUDA graph
constexpr int EDGE = 19;             // EDGE_THRESHOLD (S/ORBextractor.cc:79)
constexpr int BORDER = EDGE - 3;     // minBorderX/Y (:789)
constexpr int HALF_PATCH = 15;       // HALF_PATCH_SIZE (:78)
constexpr int MAX_DIM = 4095;        // coordinates are packed into 12 bits

struct LevelGeo {
    int w, h, pitch;                 // level size; row pitch inside the pyramid / blur slabs
    int nCols, nRows, wCell, hCell;  // FAST cell grid (:797-803)
    int cellStart;                   // first CTA index of this level in k_fast's grid
    int maxBX, maxBY;                // maxBorderX/Y (:791-792)
    int N, nIni;                     // quadtree quota and number of root nodes (:556)
    int candOff, candCap;            // slice of the per-frame candidate slab (u32 entries)
    int kpOff, kpCap;                // slice of the per-frame level-keypoint slab
    int xtabOff, ytabOff;            // resize tables (short4 entries) for producing this level
    int blurTileStart, blurTilesX;   // k_blur tile table
    float hX;                        // root node width (:558)
    float scale, kpSize;             // mvScaleFactor[l]; (int)(PATCH_SIZE*scale) (:855)
    long long pyrOff, blurOff;       // byte offsets inside a frame's slab
};

struct ExtractParams {
    int nlevels, iniTh, minTh, batch;
    const uint8_t* in;               // level 0 = the caller's frames
    long long inFrameStride;
    int inPitch, inRowBytes;         // inRowBytes: bytes of a level-0 row that may be read as whole aligned words
    int inPadded;                    // level 0 sits in the staging slab with its REFLECT_101 continuation in columns w..w+3
    uint8_t* pyr;                    // levels 1.. of every frame
    long long pyrFrameBytes;
    uint8_t* blur;                   // blurred levels 0.. of every frame
    long long blurFrameBytes;
    const short4* tabs;              // resize tables
    uint32_t* cand;                  // FAST candidates: x | y<<12 | score<<24 (x,y relative to the 16-px border)
    int* candCount;                  // batch x nlevels
    int candFrameCap;
    uint32_t* lkp;                   // per-level keypoints after the quadtree, same packing
    int* lkpCount;                   // batch x nlevels
    int kpFrameCap;
    uint16_t* qtScratch;             // node slot per candidate, used only when a tree does not fit in smem
    orbb200_keypoint* outKp;
    uint8_t* outDesc;
    int* outCount;
    int outCap;
    int* status;                     // device error bits
    int blurVariant;
    // k_fast shared-memory geometry
    int fastLarge, totalCells, totalBlurTiles;       // fastLarge: cells exceed 37 x 34 px -> the <38,64> instantiation
    const uint32_t* blurTiles;                       // k_blur tile table: level << 24 | tile row << 12 | tile column
    const int4* cells;                               // k_fast cell table (one entry per detection cell that exists, level-major)
    const int* rzXs;                                 // k_resize3: first source column (16-aligned) of every 64-column output block, level after level
    // k_resize3: a warp = 64 columns x 16 rows of two frames; per level the 16-aligned first source column of every 64-column block
    // (in rzXs from rz3XsOff on), the source-window box (rz3BoxW x rz3BoxH x 2 frames, 0: use k_resize) and the row table in groups of
    // RZ3_GROUP entries {lower source row, b0 << 16, b1 << 16, -}
    int rz3XsOff[MAXL], rz3BoxW[MAXL], rz3BoxH[MAXL], ytab3Off[MAXL];
    const int4* ytab3;
    int descChunk;                                   // k_describe3: keypoints (consecutive output rows) per group
    int nCells, frameBase;                           // entries; index of the batch's first frame inside the handle's slabs
    // k_quadtree shared-memory geometry
    int qtNC, qtPC;
    LevelGeo lv[MAXL];
};

__device__ __forceinline__ const uint8_t* level_ptr(const ExtractParams& P, int l, int frame, int& pitch)
{
    if (l == 0) { pitch = P.inPitch; return P.in + (long long)frame * P.inFrameStride; }
    pitch = P.lv[l].pitch;
    return P.pyr + (long long)frame * P.pyrFrameBytes + P.lv[l].pyrOff;
}

enum { STATUS_QT_RUNAWAY = 1, STATUS_CAND_OVERFLOW = 2, STATUS_KP_OVERFLOW = 4 };

// Level 0 in the padded staging slab: write the REFLECT_101 continuation (columns w..w+3) that k_blur reads,
// exactly as k_resize does for the other levels.  One thread per (row, frame).
__global__ void k_pad_level0(uint8_t* base, long long frameStride, int pitch, int w, int h)
{
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= h) return;
    uint8_t* row = base + (long long)blockIdx.y * frameStride + (long long)y * pitch;
#pragma unroll
    for (int i = 0; i < 4; i++) row[w + i] = row[w - 2 - i];
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Programmatic dependent launch (launch_pdl below): a kernel lets its successor in the stream start launching at once (the
// successor's blocks become resident as this grid's blocks retire and run their set-up), and waits for its predecessor's
// results right before it first touches anything an earlier kernel wrote.  Used inside the pyramid's chain of seven short
// launches (0.173 -> 0.158 ms); measured and dropped between the long kernels: behind FAST and the quadtree it changes
// nothing, and the persistent kernels (blur, descriptors) get SLOWER when their blocks become resident early (step 1.12 ->
// 1.17 ms with the blur launched this way, 1.31 ms with the descriptors too).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// ======================================================================================
// K1: pyramid level from the previous level (cv::resize INTER_LINEAR, 8UC1; :1157)
// ======================================================================================
// xtab[x] = (sx, alpha0, alpha1, sx+1 clamped), ytab[y] = (sy0, sy1, beta0, beta1), both computed on
// the host with the reference's float/double arithmetic (11-bit coefficients).
// Issue-bound, so: a thread owns 4 adjacent destination columns and walks down RZ_ROWS rows.  The
// column geometry (byte selectors, packed coefficients) is set up once; per source row the thread
// loads three aligned words, funnel-shifts them with two PRMT, and each column's horizontal blend
// S[sx]*a0 + S[sx+1]*a1 is one PRMT + one IDP.2A (coefficients in the 16-bit lanes, pixels in the
// byte lanes).  A source row shared by two consecutive destination rows is computed once.  The
// vertical blend ((b*(r>>4))>>16 twice, +2, >>2) is two IMAD.HI per pixel.
// Columns w..w+3 are written too: they hold the REFLECT_101 continuation that k_blur reads.
constexpr int RZ_ROWS = 8, RZ_WARPS = 4;

struct ResizeRaw { uint32_t w0, w1, w2; };     // the three aligned source words a thread's 4 columns draw from

__device__ __forceinline__ ResizeRaw resize_load(const uint8_t* row, int a, bool ld1, bool ld2)
{
    ResizeRaw w;
    w.w0 = __ldg(reinterpret_cast<const uint32_t*>(row + a));
    w.w1 = ld1 ? __ldg(reinterpret_cast<const uint32_t*>(row + a + 4)) : 0u;
    w.w2 = ld2 ? __ldg(reinterpret_cast<const uint32_t*>(row + a + 8)) : 0u;
    return w;
}

__device__ __forceinline__ void resize_hrow(const ResizeRaw& w, uint32_t selShift, const uint32_t (&sel)[4],
                                            const uint32_t (&coef)[4], uint32_t (&r)[4])
{
    const uint32_t lo = __byte_perm(w.w0, w.w1, selShift), hi = __byte_perm(w.w1, w.w2, selShift);
#pragma unroll
    for (int j = 0; j < 4; j++) r[j] = __dp2a_lo(coef[j], __byte_perm(lo, hi, sel[j]), 0u) >> 4;
}

__global__ void __launch_bounds__(RZ_WARPS * 32) k_resize(const ExtractParams P, int l)
{
    const LevelGeo& g = P.lv[l];
    const int frame = blockIdx.z;
    const int x0 = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4;
    const int yBeg = (blockIdx.y * RZ_WARPS + (threadIdx.x >> 5)) * RZ_ROWS;
    if (x0 >= g.w + 4 || yBeg >= g.h) return;
    int sp;
    const uint8_t* src = level_ptr(P, l - 1, frame, sp);
    const int rowBytes = (l - 1 == 0) ? P.inRowBytes : sp;    // level 0 may be the caller's buffer: never read past its pixels
    uint8_t* dst = P.pyr + (long long)frame * P.pyrFrameBytes + g.pyrOff;
    const short4* xt = P.tabs + g.xtabOff;

    // column geometry, once per thread
    short4 t[4];
    int bmin = 1 << 30;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        t[j] = __ldg(xt + min(x0 + j, g.w + 3));
        bmin = min(bmin, (int)t[j].x);
    }
    const int a = bmin & ~3;
    const uint32_t selShift = 0x3210u + 0x1111u * (uint32_t)(bmin - a);
    uint32_t sel[4], coef[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        sel[j] = (uint32_t)(t[j].x - bmin) | ((uint32_t)(t[j].w - bmin) << 4);    // bytes S[sx], S[sx+1] -> byte lanes 0,1
        coef[j] = (uint32_t)(uint16_t)t[j].y | ((uint32_t)(uint16_t)t[j].z << 16);
    }
    const bool ld1 = a + 4 < rowBytes, ld2 = a + 8 < rowBytes;

    // (the row-table entry of the next destination row is loaded one row ahead: table entry -> source row is a
    // dependent load chain; prefetching the source words as well costs more in registers than it hides, measured)
    const short4* yt = P.tabs + g.ytabOff;
    const int yEnd = min(yBeg + RZ_ROWS, g.h);
    uint32_t r0[4], r1[4];
    int cur1 = -1;                                            // source row held in r1
    short4 tnext = __ldg(yt + yBeg);
    uint8_t* outp = dst + (long long)yBeg * g.pitch + x0;
    for (int y = yBeg; y < yEnd; y++) {
        const short4 ty = tnext;
        tnext = __ldg(yt + min(y + 1, g.h - 1));
        if (ty.x == cur1) {
#pragma unroll
            for (int j = 0; j < 4; j++) r0[j] = r1[j];
        } else {
            resize_hrow(resize_load(src + (long long)ty.x * sp, a, ld1, ld2), selShift, sel, coef, r0);
        }
        if (ty.y == ty.x) {
#pragma unroll
            for (int j = 0; j < 4; j++) r1[j] = r0[j];
        } else {
            resize_hrow(resize_load(src + (long long)ty.y * sp, a, ld1, ld2), selShift, sel, coef, r1);
        }
        cur1 = ty.y;
        const uint32_t B0 = (uint32_t)ty.z << 16, B1 = (uint32_t)ty.w << 16;
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) o[j] = min((__umulhi(B0, r0[j]) + __umulhi(B1, r1[j]) + 2u) >> 2, 255u);
        *reinterpret_cast<uint32_t*>(outp) = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
        outp += g.pitch;
    }
}

struct ResizeMaps { CUtensorMap m[MAXL]; };     // m[l] = source map for producing level l (level l - 1), box rz3BoxW[l] x rz3BoxH[l] x 2 frames

// ---- k_resize, third generation: the loop runs over SOURCE rows -------------------------------------------------------------
// (The second generation, k_resize2, staged a 128 x 32 block's source window with one TMA copy and kept k_resize's row loop:
// 0.274 -> 0.236 ms.)  It still spent ~35 instructions per pixel (ncu r2s): per destination row a table load behind address arithmetic, two
// "do I already hold this source row" tests with their reconvergence points, an eight-row walk that amortises the column setup
// over 32 pixels only, and 128-column blocks that leave 16 % of the lanes past the level's right edge.  Here
//  * a warp owns 64 columns x 16 rows of TWO frames (lane = column quad + 16 * frame): the control flow depends on the row
//    tables only, so it is the same for both half-warps, and 64-column blocks fit the level widths to 94 % on average;
//  * every warp stages its own source window (one 3-D TMA box over {column, row, 2 frames}) behind its own mbarrier -- no block
//    barrier, warps of a block are independent;
//  * the loop walks the window's source rows once: each gets one horizontal pass (the two register sets swap roles), and a
//    destination row is emitted when its LOWER source row has just been computed (its upper row is then the previous one,
//    because the lower rows of consecutive destination rows never decrease).  A level with a row whose two taps sit on the same
//    source row (only when consecutive levels have the same height) keeps k_resize;
//  * the row table (a copy of the warp's 16 entries + a sentinel in shared memory) holds the lower row and the coefficients
//    already shifted (b << 16); the vertical blend packs two pixels per register before the rounding shift, and the clamp to
//    255 is gone: b0 + b1 = 2048 and r <= 255 * 2048 >> 4 bound the sum by 1020.
constexpr int RZ3_ROWS = 16, RZ3_WARPS = 4;
constexpr int RZ3_GROUP = RZ3_ROWS + 1;        // row-table entries per 16-row group: its rows, then a sentinel that matches no source row

__global__ void __launch_bounds__(RZ3_WARPS * 32) k_resize3(const __grid_constant__ ExtractParams P, const __grid_constant__ ResizeMaps M, int l)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const LevelGeo& g = P.lv[l];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int group = blockIdx.y * RZ3_WARPS + warp;
    pdl_trigger();
    if (group * RZ3_ROWS >= g.h) return;                                       // the whole warp
    const int boxW = P.rz3BoxW[l], frameBytes = boxW * P.rz3BoxH[l], warpBytes = (2 * frameBytes + 127) & ~127;
    uint8_t* win = smem + warp * warpBytes;
    int4* rows = reinterpret_cast<int4*>(smem + RZ3_WARPS * warpBytes) + warp * RZ3_GROUP;      // this warp's row-table entries
    const uint32_t bar = smem_u32(smem + RZ3_WARPS * (warpBytes + RZ3_GROUP * 16) + 8 * warp);
    const int4* yt = P.ytab3 + P.ytab3Off[l] + group * RZ3_GROUP;
    const int xs = __ldg(P.rzXs + P.rz3XsOff[l] + blockIdx.x);
    const int frame0 = blockIdx.z * 2;
    if (lane < RZ3_GROUP) rows[lane] = __ldg(yt + lane);
    if (lane == 0) {
        const int s0 = __ldg(&yt[0].x) - 1;                                    // first source row of the window
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        pdl_wait();                                                            // level l - 1 is complete
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(2 * frameBytes)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(win)), "l"(reinterpret_cast<uint64_t>(&M.m[l])), "r"(bar), "r"(xs), "r"(s0),
                       "r"((l == 1 ? 0 : P.frameBase) + frame0) : "memory");
    }
    __syncwarp();
    const int fz = lane >> 4, frame = frame0 + fz;
    const int x0 = (blockIdx.x * 16 + (lane & 15)) * 4;
    if (x0 >= g.w + 4 || frame >= P.batch) return;       // lane 0 never leaves here: the window is consumed before the block retires

    // column geometry, once per thread (as in k_resize)
    const short4* xt = P.tabs + g.xtabOff;
    short4 tx[4];
    int bmin = 1 << 30;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        tx[j] = __ldg(xt + min(x0 + j, g.w + 3));
        bmin = min(bmin, (int)tx[j].x);
    }
    const int a = bmin & ~3;
    const uint32_t selShift = 0x3210u + 0x1111u * (uint32_t)(bmin - a);
    uint32_t sel[4], coef[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        sel[j] = (uint32_t)(tx[j].x - bmin) | ((uint32_t)(tx[j].w - bmin) << 4);
        coef[j] = (uint32_t)(uint16_t)tx[j].y | ((uint32_t)(uint16_t)tx[j].z << 16);
    }
    uint32_t srow = smem_u32(win) + fz * frameBytes + (a - xs);                // source row s, column a
    uint32_t rowp = smem_u32(rows);
    uint8_t* outp = P.pyr + (long long)frame * P.pyrFrameBytes + g.pyrOff + (long long)group * RZ3_ROWS * g.pitch + x0;
    int pitch;
    asm volatile("mov.u32 %0, %1;" : "=r"(pitch) : "r"(g.pitch));              // kept in a register (otherwise re-read from the parameters per row)
    uint32_t A[4], B[4];
    int tLower;
    uint32_t tB0, tB1;
    auto next_row = [&]() {
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, _}, [%3];" : "=r"(tLower), "=r"(tB0), "=r"(tB1) : "r"(rowp));
        rowp += 16;
    };
    auto hrow = [&](uint32_t (&r)[4]) {
        uint32_t w0, w1, w2;
        asm volatile("ld.shared.u32 %0, [%3]; ld.shared.u32 %1, [%3 + 4]; ld.shared.u32 %2, [%3 + 8];" : "=r"(w0), "=r"(w1), "=r"(w2) : "r"(srow));
        const uint32_t lo = __byte_perm(w0, w1, selShift), hi = __byte_perm(w1, w2, selShift);
#pragma unroll
        for (int j = 0; j < 4; j++) r[j] = __dp2a_lo(coef[j], __byte_perm(lo, hi, sel[j]), 0u) >> 4;
        srow += boxW;
    };
    // destination rows whose lower source row is s (just computed into `cur`; `prev` holds row s - 1)
    auto flush = [&](int s, const uint32_t (&prev)[4], const uint32_t (&cur)[4]) {
        while (tLower == s) {
            uint32_t v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) v[j] = __umulhi(tB0, prev[j]) + __umulhi(tB1, cur[j]);
            const uint32_t p01 = ((v[1] << 16) + v[0] + 0x00020002u) >> 2, p23 = ((v[3] << 16) + v[2] + 0x00020002u) >> 2;
            *reinterpret_cast<uint32_t*>(outp) = __byte_perm(p01, p23, 0x6420);
            outp += pitch;
            next_row();
        }
    };
    next_row();
    int s = tLower - 1;
    mbar_wait(bar, 0);
    hrow(A);
#pragma unroll 1
    while (tLower >= 0) {                                                      // the sentinel ends the group
        s++; hrow(B); flush(s, A, B);
        if (tLower < 0) break;
        s++; hrow(A); flush(s, B, A);
    }
}

// ======================================================================================
// K2: per-cell FAST-9/16 + NMS + threshold retry (:805-849)
// ======================================================================================
// One WARP per 30-px cell.  The kernel was issue-bound and then latency-bound at low occupancy, so it
// is built for few instructions per pixel AND a small shared-memory footprint (~7 KB per warp):
//  * the cell window (+3-px halo) is staged in shared memory expanded to 16 bits per pixel, two pixels
//    per 32-bit word.  A thread works on two adjacent centre pixels at once (the two 16-bit halves of
//    a register); centre pairs sit at even tile columns, so a ring pair at an even dx is one aligned
//    LDS.32 and one at an odd dx is cut out of two neighbouring words with a single PRMT;
//  * pass 1 rejects pairs using the four opposite ring pairs (a 9-arc of one polarity contains a pixel
//    of every opposite pair) and queues the survivors by ballot, so the full test runs on dense warps;
//  * pass 2: the 16 ring differences, biased by +256 to stay positive, go through a sliding
//    min-of-9 / max-of-9 network of 3-input packed VIMNMX3.U16x2 (min3 of min3): 80 packed min/max
//    give the exact cv::cornerScore of both pixels for both polarities;
//  * pass 3: 3x3 NMS on the byte score map; the iniThFAST -> minThFAST retry is a per-cell count
//    (a corner at threshold t is exactly "score >= t", so one score map serves both thresholds).
// FAST_TPW / FAST_TH are compile-time so every ring load is `base + immediate`:
//   <26,42> covers cells up to 39 x 36 px (every level of the 640x480 ... 1920x1080 pyramids),
//   <38,64> covers the largest possible cell (57 x 57); the pitches 26 and 38 are 2 x odd (see pass 1).
constexpr int FAST_WARPS = 1;     // cells per CTA (measured: 1 -> 0.749 ms, 2 -> 0.762, 4 -> 0.779 per 256 frames); warps never synchronise with each other

template <int TPW, int TH>
struct FastGeo {
    static constexpr int SP = TPW * 2 - 4;                // score-map pitch in bytes, >= widest cell + 2, multiple of 4
    static constexpr int QCAP = (TPW - 2) * (TH - 6);     // pixel pairs per cell (incl. masked lead-in and phantom pairs), upper bound
    static constexpr int QBYTES = 2 * QCAP + 64;
    static constexpr int TILE_BYTES = TH * TPW * 4 + 64;  // (+ slack: the last pair's right neighbour word)
    static constexpr int WARP_BYTES = ((TILE_BYTES + (TH - 4) * SP + QBYTES + 15) / 16) * 16;
    // largest cell this instantiation can stage: window (cell + 6) plus up to 4 + 3 px of alignment slack
    static constexpr bool fits(int wCell, int hCell) { return ((wCell + 13) >> 2) * 2 <= TPW && hCell + 6 <= TH; }
    static_assert(TH * TPW * 4 >= 4 * (QCAP / 2 + 64), "NMS survivors are written over the dead tile");
};

#define FAST_PAIR(a, b) __byte_perm(a, b, 0x5432)     // pixels (2j+1, 2j+2) out of words j and j+1

template <int TPW, int TH>
__global__ void __launch_bounds__(FAST_WARPS * 32) k_fast(const ExtractParams P)
{
    using G = FastGeo<TPW, TH>;
    constexpr int SP = G::SP;
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr unsigned FULL = 0xffffffffu;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cell = blockIdx.x * FAST_WARPS + warp;
    if (cell >= P.totalCells) return;
    const int frame = blockIdx.y;
    int l = 0;
    while (l + 1 < P.nlevels && cell >= P.lv[l + 1].cellStart) l++;
    const LevelGeo& g = P.lv[l];
    const int c = cell - g.cellStart;
    const int ci = c / g.nCols, cj = c - ci * g.nCols;
    const int iniY = BORDER + ci * g.hCell, iniX = BORDER + cj * g.wCell;
    if (iniY >= g.maxBY - 3 || iniX >= g.maxBX - 6) return;            // :810, :819
    const int maxY = min(iniY + g.hCell + 6, g.maxBY), maxX = min(iniX + g.wCell + 6, g.maxBX);
    const int ww = maxX - iniX, wh = maxY - iniY;
    if (ww < 7 || wh < 7) return;                                       // cv::FAST finds nothing in < 7 px
    const int dw = ww - 6, dh = wh - 6;                                 // detection area (3-px FAST margin)

    uint32_t* tw = reinterpret_cast<uint32_t*>(smem + (size_t)warp * G::WARP_BYTES);   // [TH][TPW], 2 px per word
    uint8_t* score = reinterpret_cast<uint8_t*>(tw) + G::TILE_BYTES;    // [(dh+2)][SP], zero ring
    uint16_t* queue = reinterpret_cast<uint16_t*>(score + (TH - 4) * SP);
    uint32_t* klist = tw;                                               // the tile is dead by then

    int pitch;
    const uint8_t* img = level_ptr(P, l, frame, pitch);
    // 4-byte aligned load origin: tile column 0 is image column ax, the window starts at column `shift`
    // (1..4, so the first detection column is tile column 4..7 and pixel pairs can start at tile column 4)
    const int ax = (iniX - 1) & ~3, shift = iniX - ax;
    const int nwords = (shift + ww + 3) >> 2;                           // <= TPW / 2
    const uint8_t* rowp = img + (long long)iniY * pitch + ax;
    // (level 0 is 16-byte aligned too: enqueue() stages any other input into the padded slab)

    {   // stage the window (expanded to 16 bit): two rows per step when a row needs <= 16 words,
        // four independent loads in flight per lane
        const int lpr = nwords <= 16 ? 16 : 32, rpi = 32 / lpr;
        const int wl = lane & (lpr - 1), sub = lane / lpr;
        const bool colAct = wl < nwords;
        const uint8_t* q = rowp + (long long)sub * pitch + 4 * wl;
        const long long qstep = (long long)rpi * pitch;
        uint32_t* a = tw + sub * TPW + 2 * wl;
        for (int rb = 0; rb < wh; rb += 4 * rpi) {
            const int r0 = rb + sub;
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                v[u] = 0;
                if (colAct && r0 + u * rpi < wh) v[u] = __ldg(reinterpret_cast<const uint32_t*>(q + u * qstep));
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (colAct && r0 + u * rpi < wh)
                    *reinterpret_cast<uint2*>(a + u * rpi * TPW) = make_uint2(__byte_perm(v[u], 0, 0x4140), __byte_perm(v[u], 0, 0x4342));
            q += 4 * qstep;
            a += 4 * rpi * TPW;
        }
        uint4* sz = reinterpret_cast<uint4*>(score);
        for (int i = lane; i < ((dh + 2) * SP + 15) / 16; i += 32) sz[i] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();

    // Pixel pairs sit at EVEN tile words starting at tile column 4: pair i covers detection x = 2i - off and
    // 2i - off + 1, where off = shift - 1 (0..3) columns of the first pairs lie left of the area and are masked.
    // Two neighbouring pairs (a "quad") therefore start at an 8-byte aligned word: pass 1 reads them with LDS.64.
    const int off = shift - 1;
    const uint32_t* tbase = tw + 3 * TPW + 2;                            // centre pair of (row 0, pair 0)
    const int npair = (dw + off + 1) >> 1, nquad = (npair + 1) >> 1, total = nquad * dh;
    const unsigned ltmask = (1u << lane) - 1;

    // Threshold retry (:827-833) as the reference does it: the whole detection runs at iniThFAST and is repeated at
    // minThFAST only when that left the cell without a keypoint (the tile is still intact then: nothing was written to
    // klist).  Running once at min(iniTh, minTh) and filtering by score is equivalent but sends 2-3x as many pixel pairs
    // through the exact score of pass 2 (21-31 % instead of 9-18 % of the pairs survive pass 1 on the textured frames).
    // If minTh >= iniTh the second attempt cannot find anything the first did not, so it is skipped.
    int tlow = P.iniTh;
    int nk = 0;
#pragma unroll 1
    for (int attempt = 0; attempt < 2; attempt++) {
    const uint32_t T2 = (uint32_t)tlow * 0x00010001u;

    // pass 1: cheap reject on the ring pairs (0,8) (4,12) (2,10) (6,14); a lane tests one quad = 2 pairs = 4 centre
    // pixels per step from 11 shared loads.  (An odd npair makes the last quad's second pair a phantom: it reads
    // valid shared memory and pass 2 masks it by x >= dw.)
    int nq = 0;
    {
        // Lanes walk DOWN a column of quads (consecutive lanes = consecutive rows of the same quad column): with a row
        // pitch of TPW = 26 (or 38) words, sixteen consecutive rows start in sixteen different even banks, so the
        // 8-byte loads of a half-warp touch all 32 banks once.  Walking along the rows (10 quads per row, 24-word
        // pitch) cost two wavefronts per half-warp and made the kernel shared-memory bound (83 % of the data pipe).
        int k = lane / dh, r = lane - k * dh;
        const int stepK = 32 / dh, stepR = 32 - stepK * dh;
        for (int idx = lane; idx - lane < total; idx += 32) {
            bool passA, passB;
            const bool act = idx < total;
            {   // lanes past the end test quad (0, 0) and are masked afterwards: no divergent region around the loads
                const uint32_t* b = tbase + (act ? r * TPW + 2 * k : 0);
                const uint2 c = *reinterpret_cast<const uint2*>(b);                 // centres: pair A = c.x, pair B = c.y
                const uint2 lf = *reinterpret_cast<const uint2*>(b - 2), rt = *reinterpret_cast<const uint2*>(b + 2);
                const uint2 up = *reinterpret_cast<const uint2*>(b + 3 * TPW), dn = *reinterpret_cast<const uint2*>(b - 3 * TPW);
                const uint2 u2 = *reinterpret_cast<const uint2*>(b + 2 * TPW), d2 = *reinterpret_cast<const uint2*>(b - 2 * TPW);
                const uint32_t u2l = b[2 * TPW - 1], u2r = b[2 * TPW + 2], d2l = b[-2 * TPW - 1], d2r = b[-2 * TPW + 2];
                {
                    const uint32_t R4 = FAST_PAIR(c.y, rt.x), R12 = FAST_PAIR(lf.x, lf.y);
                    const uint32_t mb = __vminu2(__vimin3_u16x2(__vmaxu2(up.x, dn.x), __vmaxu2(R4, R12), __vmaxu2(u2.y, d2l)), __vmaxu2(d2.y, u2l));
                    const uint32_t md = __vmaxu2(__vimax3_u16x2(__vminu2(up.x, dn.x), __vminu2(R4, R12), __vminu2(u2.y, d2l)), __vminu2(d2.y, u2l));
                    const uint32_t hiV = c.x + T2;
                    passA = act && (__vmaxu2(mb, hiV) != hiV || __vminu2(md + T2, c.x) != c.x);
                }
                {
                    const uint32_t R4 = FAST_PAIR(rt.x, rt.y), R12 = FAST_PAIR(lf.y, c.x);
                    const uint32_t mb = __vminu2(__vimin3_u16x2(__vmaxu2(up.y, dn.y), __vmaxu2(R4, R12), __vmaxu2(u2r, d2.x)), __vmaxu2(d2r, u2.x));
                    const uint32_t md = __vmaxu2(__vimax3_u16x2(__vminu2(up.y, dn.y), __vminu2(R4, R12), __vminu2(u2r, d2.x)), __vminu2(d2r, u2.x));
                    const uint32_t hiV = c.y + T2;
                    passB = act && (__vmaxu2(mb, hiV) != hiV || __vminu2(md + T2, c.y) != c.y);
                }
            }
            const unsigned balA = __ballot_sync(FULL, passA), balB = __ballot_sync(FULL, passB);
            if (passA) queue[nq + __popc(balA & ltmask)] = (uint16_t)(r * 64 + 2 * k);
            nq += __popc(balA);
            if (passB) queue[nq + __popc(balB & ltmask)] = (uint16_t)(r * 64 + 2 * k + 1);
            nq += __popc(balB);
            k += stepK; r += stepR;
            if (r >= dh) { r -= dh; k++; }
        }
    }
    __syncwarp();
    // pass 2: exact corner score = max over the 16 arcs of 9 of min|centre - ring|, minus 1
    // (cv::cornerScore<16>).  min/max commute with the "- centre", so the sliding min-of-9 / max-of-9 network
    // runs on the raw ring values and the centre is subtracted once at the end.  Pairs with a corner are
    // compacted in place (write index <= read index) with two flag bits saying which half is a corner.
    int ncp = 0;
    for (int q0 = 0; q0 < nq; q0 += 32) {
        const int q = q0 + lane;
        bool c0 = false, c1 = false;
        int e = 0;
        if (q < nq) {
            e = queue[q];
            const int r = e >> 6, i = e & 63;
            const uint32_t* b = tbase + r * TPW + i;
            uint32_t D[16];
            {
                const uint32_t a0 = b[3 * TPW - 1], a1 = b[3 * TPW], a2 = b[3 * TPW + 1];
                D[15] = FAST_PAIR(a0, a1); D[0] = a1; D[1] = FAST_PAIR(a1, a2);
                const uint32_t z0 = b[-3 * TPW - 1], z1 = b[-3 * TPW], z2 = b[-3 * TPW + 1];
                D[9] = FAST_PAIR(z0, z1); D[8] = z1; D[7] = FAST_PAIR(z1, z2);
                D[14] = b[2 * TPW - 1]; D[2] = b[2 * TPW + 1];
                D[10] = b[-2 * TPW - 1]; D[6] = b[-2 * TPW + 1];
                D[13] = FAST_PAIR(b[TPW - 2], b[TPW - 1]); D[3] = FAST_PAIR(b[TPW + 1], b[TPW + 2]);
                D[12] = FAST_PAIR(b[-2], b[-1]); D[4] = FAST_PAIR(b[1], b[2]);
                D[11] = FAST_PAIR(b[-TPW - 2], b[-TPW - 1]); D[5] = FAST_PAIR(b[-TPW + 1], b[-TPW + 2]);
            }
            uint32_t lo3[16], hi3[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                lo3[k] = __vimin3_u16x2(D[k], D[(k + 1) & 15], D[(k + 2) & 15]);
                hi3[k] = __vimax3_u16x2(D[k], D[(k + 1) & 15], D[(k + 2) & 15]);
            }
            uint32_t lo9[16], hi9[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                lo9[k] = __vimin3_u16x2(lo3[k], lo3[(k + 3) & 15], lo3[(k + 6) & 15]);
                hi9[k] = __vimax3_u16x2(hi3[k], hi3[(k + 3) & 15], hi3[(k + 6) & 15]);
            }
            uint32_t bright = __vimax3_u16x2(lo9[0], lo9[1], lo9[2]), dark = __vimin3_u16x2(hi9[0], hi9[1], hi9[2]);
#pragma unroll
            for (int k = 3; k < 15; k += 2) {
                bright = __vimax3_u16x2(bright, lo9[k], lo9[k + 1]);
                dark = __vimin3_u16x2(dark, hi9[k], hi9[k + 1]);
            }
            bright = __vmaxu2(bright, lo9[15]);
            dark = __vminu2(dark, hi9[15]);
            // bright = max-arc-min(ring), dark = min-arc-max(ring); m = max(bright - v, v - dark), kept positive by +256
            const uint32_t V = b[0];
            const uint32_t m2 = __vmaxu2(bright + (0x01000100u - V), (0x01000100u + V) - dark);    // 256 + m per half
            const int m0 = (int)(m2 & 0xffffu) - 256, m1 = (int)(m2 >> 16) - 256;
            const int x0 = 2 * i - off;
            uint8_t* sp = score + (r + 1) * SP + x0 + 1;
            c0 = (m0 > tlow) && (x0 >= 0) && (x0 < dw);
            c1 = (m1 > tlow) && (x0 + 1 >= 0) && (x0 + 1 < dw);
            if (c0) sp[0] = (uint8_t)(m0 - 1);
            if (c1) sp[1] = (uint8_t)(m1 - 1);
        }
        const unsigned bal = __ballot_sync(FULL, c0 || c1);
        if (c0 || c1) queue[ncp + __popc(bal & ltmask)] = (uint16_t)(e | (c0 ? 0x4000 : 0) | (c1 ? 0x8000 : 0));
        ncp += __popc(bal);
    }
    __syncwarp();
    // pass 3: 3x3 non-max suppression (strictly greater than all 8 neighbours, non-corners count 0);
    // survivors go to klist, which re-uses the tile
    for (int q0 = 0; q0 < ncp; q0 += 32) {
        const int q = q0 + lane;
        const int e = q < ncp ? queue[q] : 0;
        const int r = (e >> 6) & 0xff, x0 = 2 * (e & 63) - off;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            bool keep = false;
            int s = 0;
            const int x = x0 + j;
            if (e & (0x4000 << j)) {
                const uint8_t* sp = score + (r + 1) * SP + x + 1;
                s = sp[0];
                const int nb = max(max(max((int)sp[-1], (int)sp[1]), max((int)sp[-SP - 1], (int)sp[-SP])),
                                   max(max((int)sp[-SP + 1], (int)sp[SP - 1]), max((int)sp[SP], (int)sp[SP + 1])));
                keep = s > nb;
            }
            const unsigned bal = __ballot_sync(FULL, keep);
            if (keep) {
                const int xr = x + 3 + cj * g.wCell, yr = r + 3 + ci * g.hCell;   // relative to the border (:840-841)
                klist[nk + __popc(bal & ltmask)] = (uint32_t)xr | ((uint32_t)yr << 12) | ((uint32_t)s << 24);
            }
            nk += __popc(bal);
        }
    }
    __syncwarp();
    if (nk > 0 || P.minTh >= P.iniTh) break;
    tlow = P.minTh;
    }   // attempt
    if (nk == 0) return;
    int base = 0;
    if (lane == 0) base = atomicAdd(&P.candCount[frame * P.nlevels + l], nk);
    base = __shfl_sync(FULL, base, 0);
    uint32_t* out = P.cand + (long long)frame * P.candFrameCap + g.candOff;
    for (int i = lane; i < nk; i += 32) {
        const int p = base + i;
        if (p < g.candCap) out[p] = klist[i];
        else atomicOr(P.status, STATUS_CAND_OVERFLOW);
    }
}
#undef FAST_PAIR

// ---- k_fast, second generation ---------------------------------------------------------------------------------------
// Same three passes and the same arithmetic as above; what changed is everything around them (ncu r1p: window staging was
// 16 % of the instructions and 24 % of the stall samples, cell set-up 5 %, NMS 12 %):
//  * the cell's window arrives as ONE 3-D TMA tensor copy (cp.async.bulk.tensor.3d over a {column, row, frame} map of the
//    level, SASS UTMALDG): no per-lane global addresses.  The TMA unit wants the box to start on a 16-byte boundary (measured:
//    any other column raises "illegal instruction"), so the raw bytes land up to 15 columns left of the window, in the region
//    the score map and the queue use later, and the expansion to the 16-bit tile realigns them with one funnel shift per
//    word: the window always sits at tile column 1 and the first detection pixel at tile column 4.  The pixel-pair grid
//    therefore never has masked lead-in pairs (the old kernel's `off`, up to 3 of ~17 pairs per row);
//  * cell geometry (level, window origin and size, offset of the cell inside the level) comes from a table built at create:
//    one 16-byte load instead of the level search, two divisions and the border tests;
//  * NMS handles both pixels of a pair at once on packed 16-bit lanes (3 x 4 score bytes -> nine PRMT -> four packed max);
//  * the pass-1 threshold tests are two IADD3 + one LOP3 per pair (sign bits of 0x8000 + a - b per half).
struct FastMaps { CUtensorMap m[MAXL]; };

template <int TPW, int TH>
struct FastGeo2 {
    static constexpr int SP = TPW * 2 - 4;                    // score-map pitch in bytes (multiple of 4)
    static constexpr int BW = TPW <= 26 ? 64 : 96;            // TMA box: BW bytes x TH rows (15 alignment columns + window + 5)
    static constexpr int QCAP = (TPW - 2) * (TH - 6);
    static constexpr int QBYTES = 2 * QCAP + 64;
    static constexpr int RAW_BYTES = BW * TH;
    static constexpr int REGION = (((TH - 4) * SP + QBYTES > RAW_BYTES ? (TH - 4) * SP + QBYTES : RAW_BYTES) + 127) / 128 * 128;
    static constexpr int TILE_BYTES = TH * TPW * 4 + 64;
    static constexpr int BAR_OFF = (REGION + TILE_BYTES + 15) / 16 * 16;
    static constexpr int SMEM_BYTES = BAR_OFF + 16;
    static constexpr int CHL = BW / 16 <= 4 ? 4 : 8;          // lanes per row in the expansion (>= 16-column chunks per row)
    // window (cell + 6) needs tile columns 0 .. ww + 4 and rows 0 .. wh - 1, behind up to 15 alignment columns in the box
    static constexpr bool fits(int wCell, int hCell) { return 15 + wCell + 6 + 5 <= BW && wCell + 6 + 5 <= 2 * TPW && hCell + 6 <= TH; }
    static_assert(RAW_BYTES + 32 <= REGION, "the expansion reads a whole 16-byte chunk past the last needed column");
    static_assert(TH * TPW * 4 >= 4 * (QCAP / 2 + 64), "NMS survivors are written over the dead tile");
};

#define FAST_PAIR(a, b) __byte_perm(a, b, 0x5432)

template <int TPW, int TH>
__global__ void __launch_bounds__(32) k_fast2(const __grid_constant__ ExtractParams P, const __grid_constant__ FastMaps M)
{
    using G = FastGeo2<TPW, TH>;
    constexpr int SP = G::SP, BW = G::BW;
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    const int frame = blockIdx.y;
    const int4 cr = __ldg(P.cells + blockIdx.x);
    const int iniX = cr.x & 0xffff, iniY = (int)((unsigned)cr.x >> 16);
    const int ww = cr.y & 0xff, wh = (cr.y >> 8) & 0xff, l = cr.y >> 16;
    const int dw = ww - 6, dh = wh - 6;

    uint8_t* raw = smem;                                                 // [TH][BW] bytes from the TMA unit
    uint8_t* score = smem;                                               // [(dh + 2)][SP], zero ring (after the expansion)
    uint16_t* queue = reinterpret_cast<uint16_t*>(smem + (TH - 4) * SP);
    uint32_t* tw = reinterpret_cast<uint32_t*>(smem + G::REGION);        // [TH][TPW], 2 px per word
    uint32_t* klist = tw;
    const uint32_t bar = smem_u32(smem + G::BAR_OFF);

    pdl_trigger();       // (lets the second of the two FAST launches start early; no effect otherwise)
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)G::RAW_BYTES) : "memory");
        // box origin: the 16-byte boundary at or left of image column iniX - 1 (= tile column 0; the TMA unit wants the box to
        // start on one), row iniY, frame; anything past the level's edge reads as 0
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(raw)), "l"(reinterpret_cast<uint64_t>(&M.m[l])), "r"(bar), "r"((iniX - 1) & ~15), "r"(iniY),
                       "r"((l == 0 ? 0 : P.frameBase) + frame) : "memory");
    }
    __syncwarp();
    mbar_wait(bar, 0);

    {   // expand to 16 bit and realign: tile column c = raw column c + sx (the box starts at a 16-byte boundary, sx = 0 .. 15
        // columns left of the window's tile column 0).  Lane = (row of a group, 16 tile columns); columns 0 .. ww + 4 are needed
        constexpr int RPI = 32 / G::CHL;
        const int q = lane & (G::CHL - 1), sub = lane / G::CHL;
        const int nch = (ww + 5 + 15) >> 4;
        const int sx = (iniX - 1) & 15, sh = (sx & 3) * 8;
        if (q < nch) {
            // two aligned 16-byte loads per lane (conflict-free: a quarter-warp reads 8 different 16-byte chunks), then the five
            // words from word offset sx >> 2 on -- a warp-uniform choice, so the switch costs no divergence
            const uint4* rp = reinterpret_cast<const uint4*>(raw + sub * BW) + q;
            uint32_t* d = tw + sub * TPW + 8 * q;
            const int wsel = sx >> 2;
            for (int r = sub; r < wh; r += RPI) {
                const uint4 a = rp[0], b = rp[1];
                uint32_t w0, w1, w2, w3, w4;
                switch (wsel) {
                    case 0: w0 = a.x; w1 = a.y; w2 = a.z; w3 = a.w; w4 = b.x; break;
                    case 1: w0 = a.y; w1 = a.z; w2 = a.w; w3 = b.x; w4 = b.y; break;
                    case 2: w0 = a.z; w1 = a.w; w2 = b.x; w3 = b.y; w4 = b.z; break;
                    default: w0 = a.w; w1 = b.x; w2 = b.y; w3 = b.z; w4 = b.w; break;
                }
                const uint32_t v[4] = {__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh)};
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (8 * q + 2 * k < TPW)
                        *reinterpret_cast<uint2*>(d + 2 * k) = make_uint2(__byte_perm(v[k], 0, 0x4140), __byte_perm(v[k], 0, 0x4342));
                rp += RPI * (BW / 16);
                d += RPI * TPW;
            }
        }
    }
    __syncwarp();
    {
        uint4* sz = reinterpret_cast<uint4*>(score);
        for (int i = lane; i < ((dh + 2) * SP + 15) / 16; i += 32) sz[i] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();

    // pair i covers detection x = 2i, 2i + 1 (tile columns 4 + 2i, 5 + 2i = word 2 + i); two neighbouring pairs (a quad)
    // start at an 8-byte aligned word
    const uint32_t* tbase = tw + 3 * TPW + 2;
    const int npair = (dw + 1) >> 1, nquad = (npair + 1) >> 1;
    const unsigned ltmask = (1u << lane) - 1;

    int tlow = P.iniTh;
    int nk = 0;
#pragma unroll 1
    for (int attempt = 0; attempt < 2; attempt++) {
    // per half: 0x8000 + c + t - ringmin has its top bit clear exactly when ringmin > c + t (bright candidate), and
    // 0x8000 + ringmax + t - c when ringmax < c - t (dark candidate); no borrow crosses the halves (all terms < 0x8000)
    const uint32_t TK = (uint32_t)tlow * 0x00010001u + 0x80008000u;

    int nq = 0;
    {
        // one quad (2 pairs, 4 centre pixels) per lane and step: 24 packed min/max and two sign tests.  R4a .. R12b are the
        // horizontal ring pairs (x+3, x+4) and (x-3, x-2) of the quad's two centre pairs.  `live` masks the lanes that hold a row:
        // the others compute on row 0 and are dropped from the ballots (their queue writes land behind the valid entries)
        auto eval_quad = [&](const uint2 c, const uint32_t R4a, const uint32_t R12a, const uint32_t R4b, const uint32_t R12b,
                             const uint2 up, const uint2 dn, const uint2 u2, const uint2 d2,
                             const uint32_t u2l, const uint32_t u2r, const uint32_t d2l, const uint32_t d2r, unsigned live, int entry) {
            bool passA, passB;
            {
                const uint32_t mb = __vminu2(__vimin3_u16x2(__vmaxu2(up.x, dn.x), __vmaxu2(R4a, R12a), __vmaxu2(u2.y, d2l)), __vmaxu2(d2.y, u2l));
                const uint32_t md = __vmaxu2(__vimax3_u16x2(__vminu2(up.x, dn.x), __vminu2(R4a, R12a), __vminu2(u2.y, d2l)), __vminu2(d2.y, u2l));
                passA = (((c.x + TK) - mb) & ((md + TK) - c.x) & 0x80008000u) != 0x80008000u;
            }
            {
                const uint32_t mb = __vminu2(__vimin3_u16x2(__vmaxu2(up.y, dn.y), __vmaxu2(R4b, R12b), __vmaxu2(u2r, d2.x)), __vmaxu2(d2r, u2.x));
                const uint32_t md = __vmaxu2(__vimax3_u16x2(__vminu2(up.y, dn.y), __vminu2(R4b, R12b), __vminu2(u2r, d2.x)), __vminu2(d2r, u2.x));
                passB = (((c.y + TK) - mb) & ((md + TK) - c.y) & 0x80008000u) != 0x80008000u;
            }
            const unsigned balA = __ballot_sync(FULL, passA) & live, balB = __ballot_sync(FULL, passB) & live;
            if (passA) queue[nq + __popc(balA & ltmask)] = (uint16_t)entry;
            nq += __popc(balA);
            if (passB) queue[nq + __popc(balB & ltmask)] = (uint16_t)(entry + 1);
            nq += __popc(balB);
        };
        // Rows 0 .. 31: lane = row, one quad column per step, walking right.  Consecutive lanes = consecutive rows of the same
        // column: with a row pitch of TPW = 26 (or 38) words sixteen consecutive rows start in sixteen different even banks, so
        // the 8-byte loads of a half-warp touch all 32 banks once.  The walk keeps what the next quad needs in registers: its
        // centres are this quad's right neighbours, two of its four horizontal ring pairs were cut (PRMT) for earlier quads, and
        // on rows +-2 the word pair loaded ahead for the right diagonal becomes the next quad's own pair -- five 8-byte loads
        // and two PRMT per step instead of seven loads plus four 4-byte ones and four PRMT.
        {
            const unsigned live = dh >= 32 ? FULL : (1u << dh) - 1;
            const uint32_t* b = tbase + (lane < dh ? lane * TPW : 0);
            int entry = lane * 64;
            uint2 c = *reinterpret_cast<const uint2*>(b);
            uint2 u2 = *reinterpret_cast<const uint2*>(b + 2 * TPW), d2 = *reinterpret_cast<const uint2*>(b - 2 * TPW);
            uint32_t u2l = b[2 * TPW - 1], d2l = b[-2 * TPW - 1];
            uint32_t R12a = FAST_PAIR(b[-2], b[-1]), R12b = FAST_PAIR(b[-1], c.x), Pcc = FAST_PAIR(c.x, c.y);   // Pcc: the next quad's R12a
#pragma unroll 2
            for (int k = 0; k < nquad; k++, b += 2, entry += 2) {
                const uint2 rt = *reinterpret_cast<const uint2*>(b + 2);
                const uint2 u2n = *reinterpret_cast<const uint2*>(b + 2 * TPW + 2), d2n = *reinterpret_cast<const uint2*>(b - 2 * TPW + 2);
                const uint2 up = *reinterpret_cast<const uint2*>(b + 3 * TPW), dn = *reinterpret_cast<const uint2*>(b - 3 * TPW);
                const uint32_t R4a = FAST_PAIR(c.y, rt.x), R4b = FAST_PAIR(rt.x, rt.y);
                eval_quad(c, R4a, R12a, R4b, R12b, up, dn, u2, d2, u2l, u2n.x, d2l, d2n.x, live, entry);
                R12a = Pcc; R12b = R4a; Pcc = R4b;
                c = rt; u2l = u2.y; d2l = d2.y; u2 = u2n; d2 = d2n;
            }
        }
        auto test_quad = [&](const uint32_t* b, bool act, int entry) {
            const uint2 c = *reinterpret_cast<const uint2*>(b), lf = *reinterpret_cast<const uint2*>(b - 2), rt = *reinterpret_cast<const uint2*>(b + 2);
            eval_quad(c, FAST_PAIR(c.y, rt.x), FAST_PAIR(lf.x, lf.y), FAST_PAIR(rt.x, rt.y), FAST_PAIR(lf.y, c.x),
                      *reinterpret_cast<const uint2*>(b + 3 * TPW), *reinterpret_cast<const uint2*>(b - 3 * TPW),
                      *reinterpret_cast<const uint2*>(b + 2 * TPW), *reinterpret_cast<const uint2*>(b - 2 * TPW),
                      b[2 * TPW - 1], b[2 * TPW + 2], b[-2 * TPW - 1], b[-2 * TPW + 2], __ballot_sync(FULL, act), entry);
        };
        // rows 32 .. dh - 1 (cells taller than 32 detection rows): lanes run over (row, quad column)
        const int tail = (dh - 32) * nquad;
#pragma unroll 1
        for (int idx = lane; idx - lane < tail; idx += 32) {
            const bool act = idx < tail;
            const int rr = act ? idx / nquad : 0, k = act ? idx - rr * nquad : 0;
            test_quad(tbase + (32 + rr) * TPW + 2 * k, act, (32 + rr) * 64 + 2 * k);
        }
    }
    __syncwarp();
    // pass 2: exact corner score (see k_fast)
    int ncp = 0;
    for (int q0 = 0; q0 < nq; q0 += 32) {
        const int q = q0 + lane;
        bool c0 = false, c1 = false;
        int e = 0;
        if (q < nq) {
            e = queue[q];
            const int r = e >> 6, i = e & 63;
            const uint32_t* b = tbase + r * TPW + i;
            uint32_t D[16];
            {
                const uint32_t a0 = b[3 * TPW - 1], a1 = b[3 * TPW], a2 = b[3 * TPW + 1];
                D[15] = FAST_PAIR(a0, a1); D[0] = a1; D[1] = FAST_PAIR(a1, a2);
                const uint32_t z0 = b[-3 * TPW - 1], z1 = b[-3 * TPW], z2 = b[-3 * TPW + 1];
                D[9] = FAST_PAIR(z0, z1); D[8] = z1; D[7] = FAST_PAIR(z1, z2);
                D[14] = b[2 * TPW - 1]; D[2] = b[2 * TPW + 1];
                D[10] = b[-2 * TPW - 1]; D[6] = b[-2 * TPW + 1];
                D[13] = FAST_PAIR(b[TPW - 2], b[TPW - 1]); D[3] = FAST_PAIR(b[TPW + 1], b[TPW + 2]);
                D[12] = FAST_PAIR(b[-2], b[-1]); D[4] = FAST_PAIR(b[1], b[2]);
                D[11] = FAST_PAIR(b[-TPW - 2], b[-TPW - 1]); D[5] = FAST_PAIR(b[-TPW + 1], b[-TPW + 2]);
            }
            uint32_t lo3[16], hi3[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                lo3[k] = __vimin3_u16x2(D[k], D[(k + 1) & 15], D[(k + 2) & 15]);
                hi3[k] = __vimax3_u16x2(D[k], D[(k + 1) & 15], D[(k + 2) & 15]);
            }
            uint32_t lo9[16], hi9[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                lo9[k] = __vimin3_u16x2(lo3[k], lo3[(k + 3) & 15], lo3[(k + 6) & 15]);
                hi9[k] = __vimax3_u16x2(hi3[k], hi3[(k + 3) & 15], hi3[(k + 6) & 15]);
            }
            uint32_t bright = __vimax3_u16x2(lo9[0], lo9[1], lo9[2]), dark = __vimin3_u16x2(hi9[0], hi9[1], hi9[2]);
#pragma unroll
            for (int k = 3; k < 15; k += 2) {
                bright = __vimax3_u16x2(bright, lo9[k], lo9[k + 1]);
                dark = __vimin3_u16x2(dark, hi9[k], hi9[k + 1]);
            }
            bright = __vmaxu2(bright, lo9[15]);
            dark = __vminu2(dark, hi9[15]);
            const uint32_t V = b[0];
            const uint32_t m2 = __vmaxu2(bright + (0x01000100u - V), (0x01000100u + V) - dark);    // 256 + m per half
            const int m0 = (int)(m2 & 0xffffu) - 256, m1 = (int)(m2 >> 16) - 256;
            const int x0 = 2 * i;
            c0 = (m0 > tlow) && (x0 < dw);
            c1 = (m1 > tlow) && (x0 + 1 < dw);
            // both bytes of the pair in one 16-bit store (0 = no corner; the map is zero there anyway)
            if (c0 || c1)
                *reinterpret_cast<uint16_t*>(score + (r + 1) * SP + x0 + 2) = (uint16_t)((c0 ? m0 - 1 : 0) | ((c1 ? m1 - 1 : 0) << 8));
        }
        const unsigned bal = __ballot_sync(FULL, c0 || c1);
        if (c0 || c1) queue[ncp + __popc(bal & ltmask)] = (uint16_t)e;
        ncp += __popc(bal);
    }
    __syncwarp();
    // pass 3: 3x3 NMS of both pixels of a corner pair at once.  The pair's own score bytes sit at map columns x0 + 2, x0 + 3
    // (detection x + 2, so a pair starts at an even byte); the 4 bytes x0 + 1 .. x0 + 4 of the three rows are cut out of two
    // aligned words each and spread to packed 16-bit lanes: L = (x0 - 1, x0), C = (x0, x0 + 1), R = (x0 + 1, x0 + 2).
    for (int q0 = 0; q0 < ncp; q0 += 32) {
        const int q = q0 + lane;
        const int e = q < ncp ? queue[q] : 0;
        const int r = e >> 6, x0 = 2 * (e & 63);
        bool k0 = false, k1 = false;
        uint32_t C = 0;
        if (q < ncp) {
            const uint8_t* sp = score + (r + 1) * SP + x0 + 1;
            const int sh = (int)(reinterpret_cast<uintptr_t>(sp) & 3) * 8;
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(sp - (sh >> 3));
            constexpr int SPW = SP / 4;
            const uint32_t rowU = __funnelshift_r(wp[-SPW], wp[-SPW + 1], sh);
            const uint32_t rowC = __funnelshift_r(wp[0], wp[1], sh);
            const uint32_t rowD = __funnelshift_r(wp[SPW], wp[SPW + 1], sh);
            C = __byte_perm(rowC, 0, 0x4241);
            const uint32_t nU = __vimax3_u16x2(__byte_perm(rowU, 0, 0x4140), __byte_perm(rowU, 0, 0x4241), __byte_perm(rowU, 0, 0x4342));
            const uint32_t nD = __vimax3_u16x2(__byte_perm(rowD, 0, 0x4140), __byte_perm(rowD, 0, 0x4241), __byte_perm(rowD, 0, 0x4342));
            const uint32_t nb = __vimax3_u16x2(nU, nD, __vmaxu2(__byte_perm(rowC, 0, 0x4140), __byte_perm(rowC, 0, 0x4342)));
            // keep = C > nb per half, and C != 0: a corner always has a score byte >= tlow - ... but 0 marks "no corner"
            const uint32_t gt = (nb | 0x80008000u) - C;                 // top bit of a half clear <=> C > nb
            k0 = !(gt & 0x00008000u);
            k1 = !(gt & 0x80000000u);
        }
        const unsigned bal0 = __ballot_sync(FULL, k0), bal1 = __ballot_sync(FULL, k1);
        const int xr = x0 + 3 + (cr.z & 0xffff), yr = r + 3 + (int)((unsigned)cr.z >> 16);   // relative to the border (:840-841)
        if (k0) klist[nk + __popc(bal0 & ltmask)] = (uint32_t)xr | ((uint32_t)yr << 12) | ((C & 0xffu) << 24);
        nk += __popc(bal0);
        if (k1) klist[nk + __popc(bal1 & ltmask)] = (uint32_t)(xr + 1) | ((uint32_t)yr << 12) | ((C >> 16) << 24);
        nk += __popc(bal1);
    }
    __syncwarp();
    if (nk > 0 || P.minTh >= P.iniTh) break;
    tlow = P.minTh;
    }   // attempt
    if (nk == 0) return;
    const LevelGeo& g = P.lv[l];
    int base = 0;
    if (lane == 0) base = atomicAdd(&P.candCount[frame * P.nlevels + l], nk);
    base = __shfl_sync(FULL, base, 0);
    uint32_t* out = P.cand + (long long)frame * P.candFrameCap + g.candOff;
    for (int i = lane; i < nk; i += 32) {
        const int p = base + i;
        if (p < g.candCap) out[p] = klist[i];
        else atomicOr(P.status, STATUS_CAND_OVERFLOW);
    }
}
#undef FAST_PAIR

// ======================================================================================
// K3: DistributeOctTree (:552-776) as parallel rounds
// ======================================================================================
// The reference keeps a std::list of nodes, always inserts at the front and erases split parents,
// so the list is ordered by DESCENDING creation number at all times; the output order is that
// order.  A breadth pass (:613-678) splits every node with > 1 point in list order; once
// size + 3*expandable > N (:686) nodes are split largest-first -- std::sort on (count, node address)
// read backwards (:694-699), where address order is creation order under a monotonic allocator
// (the canonical tie-break) -- stopping as soon as the list has N nodes (:743).
// Inside one round all splits are independent; the only sequential coupling is (a) the creation
// numbers of the children and (b) where the N-cut falls, and both are prefix sums over the
// processing order.  Nodes live in a slot table: the first non-empty child re-uses its parent's
// slot, so the table never needs more than max-list-size + nIni entries.
constexpr int QT_THREADS = 256;
constexpr int SIDE_MAX_BATCH = 8;      // up to this many frames per call the blur runs beside the quadtree (launch_kernels)
constexpr int QT_POINTS_ON_CHIP = 1536;   // candidates of one tree held in shared memory; larger trees run out of global memory (measured 1024 / 1536 / 2304 / 3072 / 4600: 0.125 / 0.118 / 0.118 / 0.124 / 0.140 ms per 256 VGA frames: the footprint decides how many trees share an SM)

struct QtShared {
    int size, E, cut;            // initial list (written once), position of the N-cut in a largest-first round
    int warpsum[40];
};

// Exclusive block scan of data[0 .. n) in place, returns the total.  Element i is read and written by thread i % QT_THREADS.
// `closing` = barrier after the last chunk's write-back (needed when other threads read the result next, and never for
// `warpsum`, which the callers touch again only behind later barriers).
template <bool closing>
__device__ __forceinline__ int qt_block_excl_scan(int* data, int n, int* warpsum)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int carry = 0;
    for (int base = 0; base < n; base += QT_THREADS) {
        const int i = base + tid;
        const int v = i < n ? data[i] : 0;
        int x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) warpsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            const int w = lane < QT_THREADS / 32 ? warpsum[lane] : 0;
            int s = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, s, d);
                if (lane >= d) s += y;
            }
            if (lane < QT_THREADS / 32) warpsum[lane] = s - w;
            if (lane == 31) warpsum[32] = s;
        }
        __syncthreads();
        if (i < n) data[i] = x - v + warpsum[warp] + carry;
        carry += warpsum[32];
        if (closing || base + QT_THREADS < n) __syncthreads();
    }
    return carry;
}

__device__ __forceinline__ int qt_quadrant(uint32_t p, int2 mid)
{
    const int x = p & 0xfff, y = (p >> 12) & 0xfff;
    return (x < mid.x ? 0 : 1) + (y < mid.y ? 0 : 2);
}

__global__ void __launch_bounds__(QT_THREADS) k_quadtree(const ExtractParams P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ QtShared S;
    const int tid = threadIdx.x;
    const int frame = blockIdx.y, l = blockIdx.x;
    const LevelGeo& g = P.lv[l];
    const int N = g.N, NC = P.qtNC, PC = P.qtPC;
    const int n = min(P.candCount[frame * P.nlevels + l], g.candCap);
    const uint32_t* cand = P.cand + (long long)frame * P.candFrameCap + g.candOff;
    uint32_t* lkp = P.lkp + (long long)frame * P.kpFrameCap + g.kpOff;

    // shared-memory carve-up (sizes in units of NC / PC, see extractor_create)
    unsigned long long* best = reinterpret_cast<unsigned long long*>(smem);     // NC
    int2* emid = reinterpret_cast<int2*>(best + NC);                             // NC
    short4* nrect = reinterpret_cast<short4*>(emid + NC);                        // NC (x0,x1,y0,y1)
    int* ncnt = reinterpret_cast<int*>(nrect + NC);                              // NC
    int* nseq = ncnt + NC;                                                       // NC
    int* erank = nseq + NC;                                                      // NC
    int* elist = erank + NC;                                                     // NC
    int* elist2 = elist + NC;                                                    // NC
    int* scanC = elist2 + NC;                                                    // NC
    int* scanG = scanC + NC;                                                     // NC
    int* ccnt = scanG + NC;                                                      // 4*NC
    int* cslot = ccnt + 4 * NC;                                                  // 4*NC
    uint32_t* spts = reinterpret_cast<uint32_t*>(cslot + 4 * NC);                // PC
    uint16_t* sslot = reinterpret_cast<uint16_t*>(spts + PC);                    // PC

    const uint32_t* pts;
    uint16_t* pslot;
    if (n <= PC) {
        for (int p = tid; p < n; p += QT_THREADS) spts[p] = cand[p];
        pts = spts; pslot = sslot;
    } else {   // oversized tree: run the same code out of global memory
        pts = cand;
        pslot = P.qtScratch + (long long)frame * P.candFrameCap + g.candOff;
    }
    for (int s = tid; s < NC; s += QT_THREADS) { ncnt[s] = 0; erank[s] = -1; }
    __syncthreads();

    // root nodes and initial assignment (:564-598)
    const int nIni = g.nIni;
    for (int i = tid; i < nIni; i += QT_THREADS) {
        nrect[i] = make_short4((short)(int)(g.hX * (float)i), (short)(int)(g.hX * (float)(i + 1)), 0,
                               (short)(g.maxBY - BORDER));
        nseq[i] = nIni - 1 - i;
    }
    for (int p = tid; p < n; p += QT_THREADS) {
        int r = (int)__fdiv_rn((float)(pts[p] & 0xfff), g.hX);
        r = min(max(r, 0), nIni - 1);
        pslot[p] = (uint16_t)r;
        atomicAdd(&ncnt[r], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int size = 0, E = 0;
        for (int i = 0; i < nIni; i++) {
            if (ncnt[i] > 0) size++;
            if (ncnt[i] > 1) elist[E++] = i;
        }
        S.size = size; S.E = E;
    }
    __syncthreads();

    // The state of the list (node count, slots in use, next creation number, nodes to split, phase) is carried in registers:
    // every thread derives the same next state from the same scan totals, so a round needs no bookkeeping barrier.
    int E = S.E, size = S.size, nslots = nIni, nextseq = nIni, phase = 1, rounds = 0;
    bool finish = false;
    while (!finish) {
        // A: rank the nodes to split, remember their split point
        for (int k = tid; k < E; k += QT_THREADS) {
            const int s = elist[k];
            erank[s] = k;
            const short4 r = nrect[s];
            emid[k] = make_int2(r.x + (r.y - r.x + 1) / 2, r.z + (r.w - r.z + 1) / 2);   // DivideNode :496-497
            ccnt[4 * k] = ccnt[4 * k + 1] = ccnt[4 * k + 2] = ccnt[4 * k + 3] = 0;
        }
        if (tid == 0) S.cut = E;
        __syncthreads();
        // B: count the points of each child
        for (int p = tid; p < n; p += QT_THREADS) {
            const int k = erank[pslot[p]];
            if (k >= 0) atomicAdd(&ccnt[4 * k + qt_quadrant(pts[p], emid[k])], 1);
        }
        __syncthreads();
        // C: children per node, prefix sums in processing order, N-cut
        // (a split replaces one node by its c children: the list grows by c - 1, so the running growth before node k is
        // scanC[k] - k and one scan serves both; the scan reads element k in the thread that wrote it: no barrier in between)
        for (int k = tid; k < E; k += QT_THREADS)
            scanC[k] = (ccnt[4 * k] > 0) + (ccnt[4 * k + 1] > 0) + (ccnt[4 * k + 2] > 0) + (ccnt[4 * k + 3] > 0);
        const int totalC = qt_block_excl_scan<false>(scanC, E, S.warpsum);
        if (phase == 2) {
            __syncthreads();                                                 // the cut reads its neighbour's scan element
            for (int k = tid; k < E; k += QT_THREADS)
                if (size + (k + 1 < E ? scanC[k + 1] : totalC) - (k + 1) >= N) atomicMin(&S.cut, k);        // break at :743
            __syncthreads();
        }
        const int Ecut = S.cut < E ? S.cut + 1 : E;
        const int made = Ecut < E ? scanC[Ecut] : totalC;
        const int grow = made - Ecut;
        // D: create the children (n1..n4 order = quadrant order)
        for (int k = tid; k < Ecut; k += QT_THREADS) {
            const int s = elist[k];
            const short4 r = nrect[s];
            const int2 m = emid[k];
            int j = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int cnt = ccnt[4 * k + q];
                int slot = -1;
                if (cnt > 0) {
                    slot = j == 0 ? s : nslots + scanC[k] - k + j - 1;
                    nrect[slot] = make_short4((q & 1) ? (short)m.x : r.x, (q & 1) ? r.y : (short)m.x,
                                              (q & 2) ? (short)m.y : r.z, (q & 2) ? r.w : (short)m.y);
                    ncnt[slot] = cnt;
                    nseq[slot] = nextseq + scanC[k] + j;
                    j++;
                }
                cslot[4 * k + q] = slot;
            }
        }
        __syncthreads();
        // E: move the points into their child
        for (int p = tid; p < n; p += QT_THREADS) {
            const int k = erank[pslot[p]];
            if (k >= 0 && k < Ecut) pslot[p] = (uint16_t)cslot[4 * k + qt_quadrant(pts[p], emid[k])];
        }
        // F/G: collect the children that can still be split (creation order), clear the ranks.  One scan element per split
        // node = its children with more than one point, so the scan is over Ecut elements, not 4 Ecut; nothing it writes is
        // read by step E, whose readers of erank are all past the scan's first barrier before the ranks are cleared
        for (int k = tid; k < Ecut; k += QT_THREADS)
            scanG[k] = (ccnt[4 * k] > 1) + (ccnt[4 * k + 1] > 1) + (ccnt[4 * k + 2] > 1) + (ccnt[4 * k + 3] > 1);
        const int newE = qt_block_excl_scan<false>(scanG, Ecut, S.warpsum);
        for (int k = tid; k < E; k += QT_THREADS) erank[elist[k]] = -1;
        for (int k = tid; k < Ecut; k += QT_THREADS) {
            int at = scanG[k];
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (ccnt[4 * k + q] > 1) elist2[at++] = cslot[4 * k + q];
        }
        __syncthreads();
        {
            const int nsize = size + grow;
            if (nsize >= N || nsize == size) finish = true;                              // :682 / :747
            else if (phase == 1 && nsize + 3 * newE > N) phase = 2;                      // :686
            if (++rounds > 4096) { finish = true; if (tid == 0) atomicOr(P.status, STATUS_QT_RUNAWAY); }
            if (nslots + grow > NC) { finish = true; if (tid == 0) atomicOr(P.status, STATUS_KP_OVERFLOW); }
            size = nsize; nslots += grow; nextseq += made; E = newE;
        }
        if (!finish) {
            if (phase == 1) {
                // next breadth pass walks the list front to back = newest child first (every thread writes the entry it
                // reads in the next round's step A: no barrier)
                for (int k = tid; k < newE; k += QT_THREADS) elist[k] = elist2[newE - 1 - k];
            } else {
                // largest first; equal counts: later-created node first
                for (int k = tid; k < newE; k += QT_THREADS) {
                    const int s = elist2[k], c = ncnt[s], q = nseq[s];
                    int pos = 0;
                    for (int j = 0; j < newE; j++) {
                        const int t = elist2[j];
                        const int ct = ncnt[t];
                        pos += (ct > c) || (ct == c && nseq[t] > q);
                    }
                    elist[pos] = s;
                }
                __syncthreads();
            }
        }
    }
    __syncthreads();

    // best response per leaf; ties go to the earlier candidate in the reference's order
    // (cells row-major, row-major inside a cell), recovered from the coordinates (:755-773)
    for (int s = tid; s < nslots; s += QT_THREADS) best[s] = 0ull;
    __syncthreads();
    for (int p = tid; p < n; p += QT_THREADS) {
        const uint32_t e = pts[p];
        const int x = e & 0xfff, y = (e >> 12) & 0xfff;
        const unsigned long long ord = ((unsigned long long)((y - 3) / g.hCell) << 32) |
                                       ((unsigned long long)((x - 3) / g.wCell) << 24) | ((unsigned long long)y << 12) | x;
        const unsigned long long key = ((unsigned long long)(e >> 24) << 40) | (0xffffffffffull - ord);
        atomicMax(&best[pslot[p]], key);
    }
    __syncthreads();
    // list order = descending creation number: a node's position is the number of live nodes created after it.  Creation
    // numbers are unique, so the live ones are marked in a bitmap (in ccnt, dead by now) and a position is a popcount over the
    // words above the node's bit -- a dozen words instead of a pass over all node slots (which stays as the fallback for a
    // tree that created more nodes than the bitmap holds)
    const int nwords = (nextseq + 31) >> 5;
    const bool bitmap = nwords <= 4 * NC;
    unsigned* live = reinterpret_cast<unsigned*>(ccnt);
    if (bitmap) {
        for (int i = tid; i < nwords; i += QT_THREADS) live[i] = 0u;
        __syncthreads();
        for (int s = tid; s < nslots; s += QT_THREADS)
            if (ncnt[s] > 0) atomicOr(&live[nseq[s] >> 5], 1u << (nseq[s] & 31));
        __syncthreads();
    }
    for (int s = tid; s < nslots; s += QT_THREADS) {
        if (ncnt[s] <= 0) continue;
        const int q = nseq[s];
        int pos = 0;
        if (bitmap) {
            pos = __popc(live[q >> 5] & ~(0xffffffffu >> (31 - (q & 31))));        // bits above q in its own word
            for (int w = (q >> 5) + 1; w < nwords; w++) pos += __popc(live[w]);
        } else
        for (int t = 0; t < nslots; t++) pos += (ncnt[t] > 0 && nseq[t] > q);
        const unsigned long long key = best[s];
        const unsigned long long ord = 0xffffffffffull - (key & 0xffffffffffull);
        if (pos < g.kpCap) lkp[pos] = (uint32_t)(ord & 0xffffff) | ((uint32_t)(key >> 40) << 24);
        else atomicOr(P.status, STATUS_KP_OVERFLOW);
    }
    if (tid == 0) P.lkpCount[frame * P.nlevels + l] = min(size, g.kpCap);
}

// ======================================================================================
// K5: GaussianBlur 7x7 sigma 2, BORDER_REFLECT_101 (:1117), 8.8 fixed point like OpenCV
// ======================================================================================
// A warp owns a 128 x 24 output tile.  Its 160 x 30 input window arrives as ONE 3-D TMA tensor copy (completion on an
// mbarrier) into a two-stage ring in shared memory, so the window of tile t+1 flies while tile t is being filtered and no load
// latency sits on the critical path.  (Round 1 issued one 1-D bulk copy per row: the compiler serialises such copies over the
// lanes -- ELECT, five R2UR, UBLKCP, branch -- and ncu r2z counted 13.5 % of the kernel's instructions there.)  REFLECT_101 in
// y: rows outside the level arrive as zeros and the (at most three) rows a border tile needs are copied over them inside shared
// memory; in x it is one PRMT at the left edge and (level 0 only) at the right edge -- levels >= 1 carry their reflected
// continuation in columns w..w+3 (written by k_resize).
// Arithmetic, per thread = 4 adjacent pixels: horizontal pass = ten IDP.4A of the three aligned words against constant tap
// words; vertical pass over a window of six row PAIRS of 16-bit row sums (loop unrolled by 6 so the window rotates by
// renaming), two taps per IDP.2A; out = (sum + 2^15) >> 16 is byte 2 of the accumulator.
constexpr int BL_ROWS = 24, BL_WARPS = 4;     // 24 output rows + 6 halo rows = 5 groups of 6 input rows (measured 24 / 36: 0.243 / 0.247 ms)
constexpr int BL_IN_ROWS = BL_ROWS + 6;
constexpr int BL_CTAS_PER_SM = 5;             // persistent CTAs per SM (shared memory: 9.6 KB per warp)
constexpr int BL_ROW_BYTES = 160;             // image columns x0-16 .. x0+143 (16-byte aligned window around 128 px)
constexpr int BL_WINDOW_BYTES = BL_IN_ROWS * BL_ROW_BYTES;
constexpr int BL_STAGE_BYTES = (BL_WINDOW_BYTES + 127) / 128 * 128;      // a TMA destination is 128-byte aligned
constexpr int BL_WARP_BYTES = 2 * BL_STAGE_BYTES + 128;

struct BlurMaps { CUtensorMap m[MAXL]; };     // {column, row, frame} over each level incl. its continuation columns, box 160 x 30 x 1

__device__ __forceinline__ int reflect101(int p, int n)
{
    if (p < 0) p = -p;
    if (p >= n) p = 2 * (n - 1) - p;
    return p;
}

// 18 * h + 2^15 (the newest row's tap and the rounding constant) as two shift-adds: IDP and IMAD share one half-rate pipe, which
// is what bounds this kernel (ncu r2c: the two make up 72 % of the row loop, "math pipe throttle" is the top stall reason),
// while the pipe of the shifts, adds and byte permutes is 41 % busy
__device__ __forceinline__ uint32_t blur_k0_term(uint32_t h)
{
    uint32_t t, r;
    asm("shf.l.wrap.b32 %0, 0, %1, 1;" : "=r"(t) : "r"(h));                    // 2 h  (h < 2^16)
    asm("add.u32 %0, %1, 32768;" : "=r"(t) : "r"(t));
    asm("shf.l.wrap.b32 %0, 0, %1, 4;" : "=r"(r) : "r"(h));                    // 16 h
    asm("add.u32 %0, %1, %2;" : "=r"(r) : "r"(r), "r"(t));
    return r;
}

struct BlurTile { int l, frame, x0, y0; };

// tile index -> (frame, level, position); false when the level has no keypoints (skipped like :1112).  The per-frame tile
// list is a table built at create (level << 24 | tile row << 12 | tile column).
__device__ __forceinline__ bool blur_tile(const ExtractParams& P, unsigned t, BlurTile& bt)
{
    const unsigned frame = t / (unsigned)P.totalBlurTiles, tile = t - frame * (unsigned)P.totalBlurTiles;
    const uint32_t e = __ldg(P.blurTiles + tile);
    bt.l = e >> 24; bt.frame = (int)frame; bt.x0 = (e & 0xfff) * 128; bt.y0 = ((e >> 12) & 0xfff) * BL_ROWS;
    // (a level keeps >= 1 keypoint exactly when FAST found >= 1 candidate on it: the test does not wait for the quadtree)
    return P.candCount[frame * P.nlevels + bt.l] != 0;
}

template <bool VARIANT>     // false: OpenCV >= 3 taps {18,34,48,56,48,34,18}; true: OpenCV 2.4.9 taps {18,34,49,55,49,34,18}
__global__ void __launch_bounds__(BL_WARPS * 32) k_blur(const __grid_constant__ ExtractParams P, const __grid_constant__ BlurMaps M)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* ring = smem + (size_t)warp * BL_WARP_BYTES;
    const uint32_t bar0 = smem_u32(ring + 2 * BL_STAGE_BYTES);          // two 8-byte mbarriers
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const unsigned nTiles = (unsigned)P.totalBlurTiles * (unsigned)P.batch;
    const unsigned stride = gridDim.x * BL_WARPS;
    constexpr uint32_t k0 = 18, k1 = 34, k2 = VARIANT ? 49 : 48, k3 = VARIANT ? 55 : 56;
    constexpr uint32_t kA = k0 | (k1 << 8) | (k2 << 16) | (k3 << 24);     // taps for p[x-3..x]
    constexpr uint32_t kB = k2 | (k1 << 8) | (k0 << 16);                  // taps for p[x+1..x+3]
    // the 7 taps of output column x0 + j laid over the three aligned words w0 = p[x0-4..x0-1], w1 = p[x0..x0+3], w2 = p[x0+4..x0+7]
    // (tap i sits on byte j + 1 + i of the twelve): dot products against constants, no byte shuffles
    constexpr uint32_t kT[7] = {k0, k1, k2, k3, k2, k1, k0};
    auto tapw = [](const uint32_t (&t)[7], int j, int w) constexpr {
        uint32_t v = 0;
        for (int b = 0; b < 4; b++) { const int i = 4 * w + b - j - 1; if (i >= 0 && i < 7) v |= t[i] << (8 * b); }
        return v;
    };
    constexpr uint32_t c00 = tapw(kT, 0, 0), c01 = tapw(kT, 0, 1), c10 = tapw(kT, 1, 0), c11 = tapw(kT, 1, 1), c12 = tapw(kT, 1, 2),
                       c20 = tapw(kT, 2, 0), c21 = tapw(kT, 2, 1), c22 = tapw(kT, 2, 2);
    static_assert(tapw(kT, 0, 2) == 0 && tapw(kT, 3, 0) == 0 && tapw(kT, 3, 1) == kA && tapw(kT, 3, 2) == kB, "tap layout");
    constexpr uint32_t kV01 = k0 | (k1 << 8), kV23 = k2 | (k3 << 8), kV45 = k2 | (k1 << 8);   // vertical taps by row pair

    // issue the window copy of tile bt into ring stage s: columns x0 - 16 .. x0 + 143, rows y0 - 3 .. y0 + 26 (zeros outside the level)
    auto issue = [&](const BlurTile& bt, int s) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads (and border-row writes) of this stage are done
        __syncwarp();
        if (lane == 0) {
            const uint32_t bar = bar0 + 8 * s;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)BL_WINDOW_BYTES) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(smem_u32(ring + s * BL_STAGE_BYTES)), "l"(reinterpret_cast<uint64_t>(&M.m[bt.l])), "r"(bar), "r"(bt.x0 - 16), "r"(bt.y0 - 3),
                           "r"((bt.l == 0 ? 0 : P.frameBase) + bt.frame) : "memory");
        }
    };

    unsigned t = blockIdx.x * BL_WARPS + warp;
    BlurTile cur, nxt;
    while (t < nTiles && !blur_tile(P, t, cur)) t += stride;
    if (t >= nTiles) return;
    issue(cur, 0);
    int stage = 0;
    uint32_t phase0 = 0, phase1 = 0;
    while (t < nTiles) {
        unsigned tn = t + stride;
        while (tn < nTiles && !blur_tile(P, tn, nxt)) tn += stride;
        if (tn < nTiles) issue(nxt, stage ^ 1);
        if (stage == 0) { mbar_wait(bar0, phase0); phase0 ^= 1; } else { mbar_wait(bar0 + 8, phase1); phase1 ^= 1; }

        const LevelGeo& g = P.lv[cur.l];
        {   // REFLECT_101 in y: window row ir holds image row y0 - 3 + ir; rows -3..-1 are rows 3..1 and rows h..h+2 are rows h-2..h-4
            // (rows further down only feed output rows that are not stored)
            uint32_t* st = reinterpret_cast<uint32_t*>(ring + stage * BL_STAGE_BYTES);
            bool fixed = false;
            if (cur.y0 == 0) {
#pragma unroll
                for (int r = 0; r < 3; r++)
                    for (int i = lane; i < BL_ROW_BYTES / 4; i += 32) st[r * (BL_ROW_BYTES / 4) + i] = st[(6 - r) * (BL_ROW_BYTES / 4) + i];
                fixed = true;
            }
            const int irH = g.h - (cur.y0 - 3);                                     // window row of image row h
            if (irH < BL_IN_ROWS) {
                for (int ir = irH; ir < min(irH + 3, BL_IN_ROWS); ir++) {
                    const int src = 2 * irH - 2 - ir;                               // image row 2 (h - 1) - y
                    for (int i = lane; i < BL_ROW_BYTES / 4; i += 32) st[ir * (BL_ROW_BYTES / 4) + i] = st[src * (BL_ROW_BYTES / 4) + i];
                }
                fixed = true;
            }
            if (fixed) __syncwarp();
        }
        const int x0 = cur.x0 + 4 * lane;
        if (x0 < g.w) {
            const bool left = x0 == 0, rightFix = (cur.l == 0) && !P.inPadded && (x0 + 4 >= g.w);
            const int nOut = min(BL_ROWS, g.h - cur.y0);                             // output rows of this tile
            const uint8_t* sb = ring + stage * BL_STAGE_BYTES + 12 + 4 * lane;      // word holding columns x0-4..x0-1
            uint8_t* outp = P.blur + (long long)cur.frame * P.blurFrameBytes + g.blurOff + (long long)cur.y0 * g.pitch + x0;
            // horizontal pass of one staged row: 4 sums of 7 taps (<= 255 * 256, 16 bits), ten IDP.4A against constant tap words
            auto hrow = [&](int ir, uint32_t (&h)[4]) {
                const uint32_t* rw = reinterpret_cast<const uint32_t*>(sb + ir * BL_ROW_BYTES);
                uint32_t w0 = rw[0], w1 = rw[1], w2 = rw[2];
                if (left) w0 = __byte_perm(w1, w2, 0x1234);            // p[-4..-1] = p[4], p[3], p[2], p[1]
                if (rightFix) w2 = __byte_perm(w0, w1, 0x3456);        // p[w..w+3] = p[w-2], p[w-3], p[w-4], p[w-5]
                h[0] = __dp4a(w0, c00, __dp4a(w1, c01, 0u));
                h[1] = __dp4a(w0, c10, __dp4a(w1, c11, __dp4a(w2, c12, 0u)));
                h[2] = __dp4a(w0, c20, __dp4a(w1, c21, __dp4a(w2, c22, 0u)));
                h[3] = __dp4a(w1, kA, __dp4a(w2, kB, 0u));
            };
            // Vertical pass on PAIRS of rows: Q[s][j] = H[r][j] | H[r + 1][j] << 16 for the six most recent pairs, so an
            // output row is k6 * H[newest] + three IDP.2A (two taps each) instead of seven multiplies; the six pair slots
            // rotate with the unrolled row index.
            uint32_t Q[6][4], prev[4], cur4[4];
            hrow(0, prev);
#pragma unroll
            for (int u = 1; u < 6; u++) {                                            // rows 1..5 only fill the window
                hrow(u, cur4);
#pragma unroll
                for (int j = 0; j < 4; j++) { Q[u][j] = __byte_perm(prev[j], cur4[j], 0x5410); prev[j] = cur4[j]; }
            }
#pragma unroll 1
            for (int grp = 1; grp < BL_IN_ROWS / 6; grp++) {
#pragma unroll
                for (int u = 0; u < 6; u++) {
                    const int ir = grp * 6 + u;                                      // completes output row ir - 6
                    // (rows past the tile's last output row are still filtered -- their input rows are staged, clamped to the
                    // image -- and only the store is skipped: straight-line code lets the six-row window rotate by renaming)
                    hrow(ir, cur4);
                    uint32_t acc[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        // pairs ending at rows ir-5, ir-3, ir-1 sit in slots (u+1)%6, (u+3)%6, (u+5)%6
                        acc[j] = __dp2a_lo(Q[(u + 1) % 6][j], kV01, __dp2a_lo(Q[(u + 3) % 6][j], kV23,
                                 __dp2a_lo(Q[(u + 5) % 6][j], kV45, blur_k0_term(cur4[j]))));
                        if (VARIANT) acc[j] = min(acc[j], 0x00ffffffu);          // taps sum to 257: saturate like OpenCV
                        Q[u][j] = __byte_perm(prev[j], cur4[j], 0x5410);
                        prev[j] = cur4[j];
                    }
                    const uint32_t lo = __byte_perm(acc[0], acc[1], 0x0062), hi = __byte_perm(acc[2], acc[3], 0x0062);
                    if (ir - 6 < nOut) *reinterpret_cast<uint32_t*>(outp) = __byte_perm(lo, hi, 0x5410);
                    outp += g.pitch;
                }
            }
        }
        __syncwarp();
        t = tn; cur = nxt; stage ^= 1;
    }
}

// ======================================================================================
// K4+K6: IC_Angle (:82-109) and computeOrbDescriptor (:113-152), one warp per keypoint
// ======================================================================================
// The 512 pattern points, transposed and packed at compile time: entry [k*32 + b] is point 16*b + k as
// x | y << 8.  Lane b of a warp needs point 16*b + k at step k -- 32 different addresses per access, which
// the constant cache would serialise 32-way -- so the table lives in global memory and is read through the
// read-only path (__ldg): one coalesced 64-byte L1 line per step.
struct PatternTable { short v[512]; };
constexpr signed char PATTERN_SRC[1024] = {
#include "../../include/orb_b200_pattern.inc"
};
constexpr PatternTable make_pattern_table()
{
    PatternTable t{};
    for (int i = 0; i < 512; i++) {
        const int byte = i >> 4, k = i & 15;
        t.v[k * 32 + byte] = (short)((PATTERN_SRC[2 * i] & 0xff) | ((int)PATTERN_SRC[2 * i + 1] * 256));
    }
    return t;
}
__device__ const PatternTable d_pattern = make_pattern_table();

// IC_Angle weights.  The radius-15 disc is read as 8 aligned-shifted words per row (columns u = -16 .. 15), four rows
// per warp step: in step `it` lane L owns word j = L & 7 of row v = -15 + 4*it + (L >> 3).  Its two weight words hold,
// per byte, u (inside the disc, else 0) and v (inside, else 0), so the moments are two IDP.4A per lane and step.
struct AngleTable { uint2 w[8][32]; };
constexpr int UMAX_SRC[HALF_PATCH + 1] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
constexpr AngleTable make_angle_table()
{
    AngleTable t{};
    for (int it = 0; it < 8; it++)
        for (int L = 0; L < 32; L++) {
            const int v = -HALF_PATCH + 4 * it + (L >> 3), j = L & 7;
            unsigned wu = 0, wv = 0;
            for (int b = 0; b < 4; b++) {
                const int u = -16 + 4 * j + b;
                const int au = u < 0 ? -u : u, av = v < 0 ? -v : v;
                if (av <= HALF_PATCH && au <= HALF_PATCH && au <= UMAX_SRC[av]) {
                    wu |= (unsigned)(u & 0xff) << (8 * b);
                    wv |= (unsigned)(v & 0xff) << (8 * b);
                }
            }
            t.w[it][L].x = wu; t.w[it][L].y = wv;
        }
    return t;
}
__device__ const AngleTable d_angle = make_angle_table();

// The same weights for k_describe3, whose patch rows sit 48 bytes = 12 words apart in shared memory: with four CONSECUTIVE rows
// per step the four 8-word windows start at banks 0, 12, 24, 4 and overlap (two wavefronts per load); with rows 2 apart they
// start at 0, 24, 16, 8 -- every bank once, also for the neighbouring word the funnel shift needs.  Step `it` takes rows
// v = -15 + 8 * (it >> 1) + (it & 1) + 2 * (L >> 3).
__host__ __device__ constexpr int angle3_row(int it, int grp) { return 8 * (it >> 1) + (it & 1) + 2 * grp; }      // patch row 0 .. 31 (31: outside, weights 0)
constexpr AngleTable make_angle_table3()
{
    AngleTable t{};
    for (int it = 0; it < 8; it++)
        for (int L = 0; L < 32; L++) {
            const int v = -HALF_PATCH + angle3_row(it, L >> 3), j = L & 7;
            unsigned wu = 0, wv = 0;
            for (int b = 0; b < 4; b++) {
                const int u = -16 + 4 * j + b;
                const int au = u < 0 ? -u : u, av = v < 0 ? -v : v;
                if (av <= HALF_PATCH && au <= HALF_PATCH && au <= UMAX_SRC[av]) {
                    wu |= (unsigned)(u & 0xff) << (8 * b);
                    wv |= (unsigned)(v & 0xff) << (8 * b);
                }
            }
            t.w[it][L].x = wu; t.w[it][L].y = wv;
        }
    return t;
}
__device__ const AngleTable d_angle3 = make_angle_table3();

__device__ __forceinline__ int dp4a_u8s8(uint32_t pixels, uint32_t weights, int acc)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(pixels), "r"(weights), "r"(acc));
    return d;
}

// cv::fastAtan2: every step is a separately rounded fp32 operation (no FMA contraction).
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float scale = (float)(180.0 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float eps = 2.2204460492503131e-16f;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// glibc sinf/cosf (ARM optimized-routines sincosf): pi/2 reduction and polynomial in double,
// one rounding to float.  Bit-identical to libm on every float in [0, 2*pi] (tested exhaustively
// on the host restatement, oracle/orb_oracle.c:orc_sincosf).
__device__ __forceinline__ float sc_poly(double x, double x2, int n, bool negcos)
{
    const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10,
                 C4 = 0x1.99343027bf8c3p-16;
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        const double x3 = __dmul_rn(x, x2);
        const double s1 = __dadd_rn(S2, __dmul_rn(x2, S3));
        const double x7 = __dmul_rn(x3, x2);
        const double s = __dadd_rn(x, __dmul_rn(x3, S1));
        return __double2float_rn(__dadd_rn(s, __dmul_rn(x7, s1)));
    }
    const double sg = negcos ? -1.0 : 1.0;
    const double x4 = __dmul_rn(x2, x2);
    const double c2 = __dadd_rn(sg * C3, __dmul_rn(x2, sg * C4));
    const double c1 = __dadd_rn(sg * C0, __dmul_rn(x2, sg * C1));
    const double x6 = __dmul_rn(x4, x2);
    const double c = __dadd_rn(c1, __dmul_rn(x4, sg * C2));
    return __double2float_rn(__dadd_rn(c, __dmul_rn(x6, c2)));
}

__device__ __forceinline__ void libm_sincosf(float y, float& s, float& c)
{
    const double HPI_INV = 0x1.45F306DC9C883p+23, HPI = 0x1.921FB54442D18p0;
    double x = (double)y;
    const uint32_t top = (__float_as_uint(y) >> 20) & 0x7ff;
    if (top < 0x3f4) {
        if (top < 0x398) { s = y; c = 1.0f; return; }
        const double x2 = __dmul_rn(x, x);
        s = sc_poly(x, x2, 0, false);
        c = sc_poly(x, x2, 1, false);
        return;
    }
    const double r = __dmul_rn(x, HPI_INV);
    const int n = ((int)r + 0x800000) >> 24;
    x = __dsub_rn(x, __dmul_rn((double)n, HPI));
    const double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    const bool neg = (n & 2) != 0;
    const double xs = __dmul_rn(x, sgn), x2 = __dmul_rn(x, x);
    s = sc_poly(xs, x2, n, neg);
    c = sc_poly(xs, x2, n ^ 1, neg);
}

constexpr int DESC_WARPS = 8;

__global__ void __launch_bounds__(DESC_WARPS * 32) k_describe(const ExtractParams P)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.y;
    const int idx = blockIdx.x * DESC_WARPS + warp;        // slot in the frame's level-keypoint slab
    if (idx >= P.kpFrameCap) return;
    int l = 0;
    while (l + 1 < P.nlevels && idx >= P.lv[l + 1].kpOff) l++;
    const LevelGeo& g = P.lv[l];
    const int i = idx - g.kpOff;
    // lane q holds level q's keypoint count: the slot's output offset and the frame total are two warp sums
    const int cnt = lane < P.nlevels ? __ldg(P.lkpCount + frame * P.nlevels + lane) : 0;
    const int off = __reduce_add_sync(0xffffffffu, lane < l ? cnt : 0);
    if (idx == 0) {                             // slot 0 publishes the frame total (also when level 0 is empty)
        const int tot = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) P.outCount[frame] = tot;
    }
    if (i >= __shfl_sync(0xffffffffu, cnt, l)) return;
    const int o = off + i;
    if (o >= P.outCap) { if (lane == 0) atomicOr(P.status, STATUS_KP_OVERFLOW); return; }

    const uint32_t e = P.lkp[(long long)frame * P.kpFrameCap + idx];
    const int x = (e & 0xfff) + BORDER, y = ((e >> 12) & 0xfff) + BORDER, score = e >> 24;   // :857-858

    // ---- orientation: intensity centroid over the radius-15 disc of the UNBLURRED level ----
    int pitch;
    const uint8_t* c = level_ptr(P, l, frame, pitch);
    c += (long long)y * pitch + x;
    int m10 = 0, m01 = 0;
    {
        // aligned words around columns x-16 .. x+15 (the level's pitch is a multiple of 4, so one shift serves all rows)
        const uint8_t* row0 = c - 16 + (long long)(-HALF_PATCH + (lane >> 3)) * pitch;
        const int sh = (int)(reinterpret_cast<uintptr_t>(row0) & 3);
        const uint32_t* p = reinterpret_cast<const uint32_t*>(row0 - sh) + (lane & 7);
        const long long step = (long long)pitch;               // 4 rows, in words
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const uint2 w = __ldg(&d_angle.w[it][lane]);
            if (it < 7 || lane < 24) {                             // the last step holds rows 13, 14, 15 only
                const uint32_t px = __funnelshift_r(__ldg(p), __ldg(p + 1), 8 * sh);
                m10 = dp4a_u8s8(px, w.x, m10);
                m01 = dp4a_u8s8(px, w.y, m01);
            }
            p += step;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, d);
        m01 += __shfl_xor_sync(0xffffffffu, m01, d);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- descriptor on the BLURRED level: lane b produces byte b (pairs 8b..8b+7) ----
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.0);
    float a, b;
    libm_sincosf(__fmul_rn(angle, factorPI), b, a);
    const uint8_t* cb = P.blur + (long long)frame * P.blurFrameBytes + g.blurOff + (long long)y * g.pitch + x;
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int t[2];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const int pk = __ldg(&d_pattern.v[(2 * k + s) * 32 + lane]);
            const float px = (float)(signed char)(pk & 0xff), py = (float)(pk >> 8);
            const int ry = __float2int_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)));
            const int rx = __float2int_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)));
            t[s] = __ldg(cb + ry * g.pitch + rx);
        }
        val |= (t[0] < t[1]) << k;
    }
    P.outDesc[((long long)frame * P.outCap + o) * 32 + lane] = (uint8_t)val;

    // ---- keypoint record (cv::KeyPoint layout), coordinates scaled to level 0 (:1126-1132) ----
    if (lane < 7) {
        float fx = (float)x, fy = (float)y;
        if (l != 0) { fx = __fmul_rn(fx, g.scale); fy = __fmul_rn(fy, g.scale); }
        uint32_t w;
        switch (lane) {
            case 0: w = __float_as_uint(fx); break;
            case 1: w = __float_as_uint(fy); break;
            case 2: w = __float_as_uint(g.kpSize); break;
            case 3: w = __float_as_uint(angle); break;
            case 4: w = __float_as_uint((float)score); break;
            case 5: w = (uint32_t)l; break;
            default: w = 0xffffffffu; break;
        }
        reinterpret_cast<uint32_t*>(P.outKp + (long long)frame * P.outCap + o)[lane] = w;
    }
}

// ---- k_describe, second generation (k_describe2, superseded by k_describe3 below, which keeps its data path) -------------
// Same arithmetic, different data path (ncu r2i: the first version was bound by L1 wavefronts, 79 %, because every one of the
// 16 byte gathers of a lane is its own sector, and by the conversion pipe, 58 %, for int -> float of the pattern and
// float -> int of the rotated coordinates; r2k: a one-keypoint-per-warp TMA version waited on its own copy, long scoreboard):
//  * persistent warps (3 CTAs per SM), each walks one contiguous range of the batch's (frame, output row) slots.  What depends on the lane only
//    -- its 16 pattern points as floats (no I2F) and its 8 IC_Angle weight words -- is loaded once and stays in registers;
//  * the keypoint's two patches -- the unblurred level around it for IC_Angle (48 bytes x 31 rows), the blurred one for BRIEF
//    (64 x 37) -- arrive as two 3-D TMA tensor copies (box start on the 16-byte boundary left of the patch) into one of the
//    warp's two shared-memory buffers, and the copies of keypoint n + 1 are issued before keypoint n is worked on;
//  * cvRound is the magic-constant addition (1.5 * 2^23, round-to-nearest-even like cvRound, exact because the value is
//    already a float far below 2^22): no F2I either.
struct DescMaps { CUtensorMap u[MAXL], b[MAXL]; };      // unblurred levels: box 48 x 31 x 1; blurred levels: box 64 x 37 x 1
constexpr int DESC_UW = 48, DESC_UROWS = 2 * HALF_PATCH + 1, DESC_BW = 64, DESC_BROWS = 37, DESC_HALF = 18;
constexpr int DESC_USLOT = (DESC_UW * DESC_UROWS + 127) / 128 * 128, DESC_BSLOT = (DESC_BW * DESC_BROWS + 127) / 128 * 128;
constexpr int DESC_BUF = DESC_USLOT + DESC_BSLOT;        // one buffer = both patches of one keypoint (TMA destinations: 128-byte aligned)
constexpr int DESC_CTAS_PER_SM = 3;                     // persistent CTAs per SM (registers: 80 x 256; shared memory: 63 KB)

struct PatternTableF { float2 v[512]; };
constexpr PatternTableF make_pattern_table_f()
{
    PatternTableF t{};
    for (int i = 0; i < 512; i++) {
        const int byte = i >> 4, k = i & 15;
        t.v[k * 32 + byte].x = (float)PATTERN_SRC[2 * i];
        t.v[k * 32 + byte].y = (float)PATTERN_SRC[2 * i + 1];
    }
    return t;
}
__device__ const PatternTableF d_pattern_f = make_pattern_table_f();

// ---- k_describe, third generation: the per-keypoint scalar work is done for a group of keypoints at once ---------------------
// ncu (r2z) on k_describe2: the moments and the 256 comparisons are only ~40 % of its 624 warp instructions per keypoint; the rest
// is work every lane repeats with the same operands -- finding the slot's frame / level / keypoint (12 %, an integer division in
// it), fastAtan2 and the double-precision sin/cos (16 %), the cv::KeyPoint record (7 %), issuing the copies (8 %).  Here a warp
// takes GROUPS of P.descChunk (8; 1 in small calls) consecutive output rows of one frame, and works on two groups at a time:
//  * lane i finds keypoint i of a group (level by comparing with the frame's running level counts, one load of the packed level
//    keypoint): once per group;
//  * a step = the moments of keypoint k of group g (unblurred patch) AND the descriptor of keypoint k of group g - 1 (blurred
//    patch, sin/cos by shuffle from lane k).  Both patches arrive as TMA copies behind one mbarrier, the next step's copies are in
//    flight while this one is worked on -- the same latency cover as k_describe2, where a step was one keypoint's two patches;
//  * between two rounds the lanes compute the angles, sin/cos and keypoint records of group g side by side: one pass per group.
// (First attempt, measured: moments of a whole group, then its descriptors, one ring of patch slots with two copies ahead --
// 0.245 ms against k_describe2's 0.209: a moments step is too short for two copies ahead to cover the copy latency.)
constexpr int D3_WARP_BYTES = 2 * DESC_BUF + 128;          // two step buffers (unblurred + blurred patch) + the warp's two mbarriers
constexpr int D3_CTA_BYTES = DESC_WARPS * D3_WARP_BYTES + 8 * 32 * 8;      // + the IC_Angle weight words

__global__ void __launch_bounds__(DESC_WARPS * 32, 3) k_describe3(const __grid_constant__ ExtractParams P, const __grid_constant__ DescMaps M)
{
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int DG = P.descChunk;                                                // keypoints per group (<= 32)
    const int gpf = (P.outCap + DG - 1) / DG;                                  // groups per frame
    const int nItems = P.batch * gpf, nw = gridDim.x * DESC_WARPS;
    int item = blockIdx.x * DESC_WARPS + warp;

    // per-lane constants: level tables and the lane's 16 pattern points
    const int kpOffLane = lane < P.nlevels ? P.lv[lane].kpOff : 0;
    const float scaleLane = lane < P.nlevels ? P.lv[lane].scale : 0.f, sizeLane = lane < P.nlevels ? P.lv[lane].kpSize : 0.f;
    float2 pat[16];
#pragma unroll
    for (int k = 0; k < 16; k++) pat[k] = __ldg(&d_pattern_f.v[k * 32 + lane]);
    // (the 8 IC_Angle weight words of every lane sit in shared memory, one conflict-free 8-byte load per step: in registers they
    // pushed the kernel past the 80 it may use at three blocks per SM)
    const uint2* wgt = reinterpret_cast<const uint2*>(smem + DESC_WARPS * D3_WARP_BYTES);
    for (int i = threadIdx.x; i < 8 * 32; i += DESC_WARPS * 32) const_cast<uint2*>(wgt)[i] = __ldg(&d_angle3.w[i >> 5][i & 31]);
    __syncthreads();
    wgt += lane;

    uint8_t* buf = smem + (size_t)warp * D3_WARP_BYTES;
    const uint32_t buf32 = smem_u32(buf), bar0 = buf32 + 2 * DESC_BUF;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (item >= nItems) return;

    // A group: frame, first output row, number of keypoints (0: no group), and in lane i keypoint i's packed level record and level.
    int nN = 0, frameN = 0, o0N = 0, lN = 0;  uint32_t eN = 0;                 // the group after next (decoded ahead)
    int nA = 0, frameA = 0, o0A = 0, lA = 0;  uint32_t eA = 0;                 // the group whose moments are being taken
    int nD = 0, frameD = 0, o0D = 0;          uint32_t eD = 0;                 // the group whose descriptors are being made
    float caD = 0.f, sbD = 0.f;                                                // ... and lane i's cos / sin of keypoint i's angle
    // the next non-empty group of this warp at or after `item` into N (nN = 0: none left)
    auto decode = [&]() {
        nN = 0;
        while (item < nItems) {
            const int f = item / gpf, o0 = (item - f * gpf) * DG;
            item += nw;
            const int cnt = lane < P.nlevels ? __ldg(P.lkpCount + f * P.nlevels + lane) : 0;
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 16; d <<= 1) { const int v = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += v; }   // MAXL = 16 levels
            const int total = min(__shfl_sync(FULL, incl, 15), P.outCap);     // (more cannot happen: outCap = the sum of the level capacities)
            if (o0 == 0 && lane == 0) P.outCount[f] = total;                  // the frame's first group publishes its total (also 0)
            const int n = min(total - o0, DG);
            if (n <= 0) continue;
            const int o = o0 + lane;
            int l = 0;
            for (int q = 0; q + 1 < P.nlevels; q++) l += (__shfl_sync(FULL, incl, q) <= o) ? 1 : 0;
            const int pre = __shfl_sync(FULL, incl - cnt, l), off = __shfl_sync(FULL, kpOffLane, l);
            frameN = f; o0N = o0; nN = n; lN = l;
            eN = lane < n ? __ldg(P.lkp + (long long)f * P.kpFrameCap + off + (o - pre)) : 0u;
            return;
        }
    };
    // the copies of one step into buffer s: the unblurred patch of keypoint k of group (na, ...) and the blurred patch of
    // keypoint k of group (nd, ...), whichever exist (at least one does)
    auto issue = [&](int na, int fa, uint32_t ea, int la, int nd, int fd, uint32_t ed, int ld, int k, int s) {
        const uint32_t ka = __shfl_sync(FULL, ea, k), kd = __shfl_sync(FULL, ed, k);
        const int lka = __shfl_sync(FULL, la, k), lkd = __shfl_sync(FULL, ld, k);
        const uint32_t bar = bar0 + 8 * s, dst = buf32 + s * DESC_BUF;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // earlier generic reads of this buffer are done
        __syncwarp();
        if (lane == 0) {
            const uint32_t bytes = (k < na ? DESC_UW * DESC_UROWS : 0) + (k < nd ? DESC_BW * DESC_BROWS : 0);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            if (k < na) {
                const int x = (ka & 0xfff) + BORDER, y = ((ka >> 12) & 0xfff) + BORDER;
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&M.u[lka])), "r"(bar), "r"((x - 16) & ~15), "r"(y - HALF_PATCH),
                               "r"((lka == 0 ? 0 : P.frameBase) + fa) : "memory");
            }
            if (k < nd) {
                const int x = (kd & 0xfff) + BORDER, y = ((kd >> 12) & 0xfff) + BORDER;
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(dst + DESC_USLOT), "l"(reinterpret_cast<uint64_t>(&M.b[lkd])), "r"(bar), "r"((x - DESC_HALF) & ~15), "r"(y - DESC_HALF),
                               "r"(P.frameBase + fd) : "memory");
            }
        }
    };

    decode();
    if (nN == 0) return;
    nA = nN; frameA = frameN; o0A = o0N; lA = lN; eA = eN;
    decode();
    int lD = 0;
    issue(nA, frameA, eA, lA, 0, 0, 0u, 0, 0, 0);
    uint32_t phase = 0;                                                       // bit s = parity to wait for on buffer s
    int s = 0;
#pragma unroll 1
    while (nA | nD) {
        const int K = max(nA, nD);
        const int xA = (eA & 0xfff) + BORDER, xD = (eD & 0xfff) + BORDER;     // lane i: keypoint i of either group
        uint8_t* outd = P.outDesc + ((long long)frameD * P.outCap + o0D) * 32 + lane;
        int M10 = 0, M01 = 0;
#pragma unroll 1
        for (int k = 0; k < K; k++) {
            // the next step's copies: this round's step k + 1, or step 0 of the next round (moments of N, descriptors of A)
            if (k + 1 < K) issue(nA, frameA, eA, lA, nD, frameD, eD, lD, k + 1, s ^ 1);
            else if (nN | nA) issue(nN, frameN, eN, lN, nA, frameA, eA, lA, 0, s ^ 1);
            mbar_wait(bar0 + 8 * s, (phase >> s) & 1);
            phase ^= 1u << s;
            const uint8_t* pu = buf + s * DESC_BUF;                           // unblurred patch: rows y - 15 .. y + 15, columns from (x - 16) & ~15
            if (k < nA) {
                // ---- intensity centroid over the radius-15 disc of the UNBLURRED level (weights: see d_angle) ----
                const int x = __shfl_sync(FULL, xA, k);
                int m10 = 0, m01 = 0;
                // columns x - 16 .. x + 15 = patch bytes b0 .. b0 + 31; lane owns word j = lane & 7 of patch row angle3_row(it, lane >> 3)
                const int b0 = (x - 16) & 15, sh = (b0 & 3) * 8;
                const uint32_t* p = reinterpret_cast<const uint32_t*>(pu + 2 * (lane >> 3) * DESC_UW) + (b0 >> 2) + (lane & 7);
#pragma unroll
                for (int it = 0; it < 8; it++) {
                    if (it < 7 || lane < 24) {                                 // the last step holds rows 10, 12, 14 only
                        const uint32_t* q = p + angle3_row(it, 0) * (DESC_UW / 4);
                        const uint32_t px = __funnelshift_r(q[0], q[1], sh);
                        const uint2 w = wgt[it * 32];
                        m10 = dp4a_u8s8(px, w.x, m10);
                        m01 = dp4a_u8s8(px, w.y, m01);
                    }
                }
                m10 = __reduce_add_sync(FULL, m10);
                m01 = __reduce_add_sync(FULL, m01);
                if (lane == k) { M10 = m10; M01 = m01; }
            }
            if (k < nD) {
                // ---- descriptor on the BLURRED level: lane b produces byte b (pairs 8b..8b+7) ----
                const int x = __shfl_sync(FULL, xD, k);
                const float a = __shfl_sync(FULL, caD, k), b = __shfl_sync(FULL, sbD, k);
                // cvRound(v) = bits(v + 1.5 * 2^23) - 0x4B400000 (round to nearest even, |v| < 2^22).  The sample's shared-memory address
                // is base + ry * 64 + rx in 32-bit arithmetic that wraps, so the two constants are folded into the base once
                const float MAGIC = 12582912.0f;
                const uint32_t cb = smem_u32(pu + DESC_USLOT) + DESC_HALF * DESC_BW + ((x - DESC_HALF) & 15) + DESC_HALF - 0x4B400000u * (uint32_t)(DESC_BW + 1);
                int val = 0;
#pragma unroll
                for (int kk = 0; kk < 8; kk++) {
                    uint32_t t[2];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const float2 pt = pat[2 * kk + h];
                        const uint32_t ry = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(pt.x, b), __fmul_rn(pt.y, a)), MAGIC));
                        const uint32_t rx = __float_as_uint(__fadd_rn(__fsub_rn(__fmul_rn(pt.x, a), __fmul_rn(pt.y, b)), MAGIC));
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t[h]) : "r"(cb + ry * (uint32_t)DESC_BW + rx));
                    }
                    val |= (t[0] < t[1]) << kk;
                }
                outd[k * 32] = (uint8_t)val;
            }
            s ^= 1;
        }
        // ---- group A's angles, sin/cos and keypoint records, lane i for keypoint i; A becomes D, N becomes A ----
        const float angle = fast_atan2_deg((float)M01, (float)M10);
        const float factorPI = (float)(3.1415926535897932384626433832795 / 180.0);
        libm_sincosf(__fmul_rn(angle, factorPI), sbD, caD);
        const float sc = __shfl_sync(FULL, scaleLane, lA), ksz = __shfl_sync(FULL, sizeLane, lA);
        if (lane < nA) {                                                       // cv::KeyPoint layout, coordinates scaled to level 0 (:1126-1132)
            float fx = (float)xA, fy = (float)(int)(((eA >> 12) & 0xfff) + BORDER);
            if (lA != 0) { fx = __fmul_rn(fx, sc); fy = __fmul_rn(fy, sc); }
            uint32_t* rec = reinterpret_cast<uint32_t*>(P.outKp + (long long)frameA * P.outCap + o0A + lane);
            rec[0] = __float_as_uint(fx); rec[1] = __float_as_uint(fy); rec[2] = __float_as_uint(ksz); rec[3] = __float_as_uint(angle);
            rec[4] = __float_as_uint((float)(eA >> 24)); rec[5] = (uint32_t)lA; rec[6] = 0xffffffffu;
        }
        __syncwarp();
        nD = nA; frameD = frameA; o0D = o0A; eD = eA; lD = lA;
        nA = nN; frameA = frameN; o0A = o0N; eA = eN; lA = lN;
        if (nA) decode(); else nN = 0;
    }
}

}  // namespace orbb200

// ======================================================================================
// host side: handle, geometry, launches, C ABI
// ======================================================================================
using namespace orbb200;

// Launch with programmatic stream serialization: the kernel may start once its predecessor in the stream has let it
// (pdl_trigger), and orders itself behind the predecessor's results with pdl_wait.  Behind anything that is not a kernel
// (a memset, an event) the attribute has no effect; the kernel behind such a launch needs no change (FAST after the pyramid's
// last level waits for the whole grid as usual).
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

static size_t resize3_smem(const ExtractParams& P, int l)
{
    return (size_t)RZ3_WARPS * ((((size_t)2 * P.rz3BoxW[l] * P.rz3BoxH[l] + 127) & ~(size_t)127) + RZ3_GROUP * 16 + 8);
}

struct orbb200_extractor {
    int nfeatures, nlevels, iniTh, minTh, width, height, maxBatch, device, blurTaps;
    double scaleFactorD;
    float scale[MAXL], invScale[MAXL], sigma2[MAXL], invSigma2[MAXL];
    int perLevel[MAXL];
    ExtractParams P;           // device pointers + geometry (in / batch patched per call)
    cudaStream_t stream;
    uint8_t* dIn;              // staging for host frames (level 0)
    size_t inPitch;
    int totalCells, totalBlurTiles, maxKp, numSMs;
    size_t fastSmem, qtSmem;
    int fastVariant;           // 2 = k_fast2 (TMA tensor staging, default), 1 = k_fast (ORBB200_FAST_VARIANT=1: the first-generation kernel, kept for A/B runs)
    ResizeMaps resizeMaps;     // k_resize3's source-window maps (m[1] is encoded per call: level 0 may be the caller's buffer)
    int resizeVariant;         // 3 = k_resize3 (source-row walk over TMA-staged windows, default), 1 = k_resize (ORBB200_RESIZE_VARIANT=1 at create, kept for A/B runs)
    BlurMaps blurMaps;         // k_blur's window maps (level 0 per call)
    int fastSmallCells;        // k_fast2: the first fastSmallCells entries of the cell table are at most 33 wide and run in the <22,42> instantiation
                               // (6.7 KB of shared memory per cell instead of 8.0: 29 cells per SM instead of 25), the rest in <26,42>
    FastMaps fastMaps;         // tensor maps of the pyramid levels (level 0 is encoded per call: it may be the caller's buffer)
    DescMaps descMaps;         // k_describe2's patch maps of the unblurred and the blurred levels (box 64 x 37)
    int descVariant, maxLevelKpCap;   // 3 = k_describe3 (default), 1 = k_describe (ORBB200_DESCRIBE_VARIANT=1 at create, kept for A/B runs)
    int lastLaunches, lastBatch;
    int realCandCap[MAXL];     // per-level candidate capacity as computed at create (orbb200_extractor_debug_set_capacity clamps P.lv[l].candCap)
    int chunkOverride;         // ORBB200_CHUNKS read once at create (0 = choose by batch size): tuning knob of the blocking host call
    const uint8_t* lastIn; long long lastInFrameStride; int lastInPitch;
    orbb200_keypoint* dOutKp; uint8_t* dOutDesc; int* dOutCount;
    void* pinned; size_t pinnedBytes;
    bool profiling; cudaEvent_t ev[6];   // stage boundaries of the last call: resize | fast | quadtree | blur | describe
    cudaStream_t copyIn, copyOut;        // host path: H2D and D2H run beside the kernels, chunk by chunk
    cudaEvent_t evIn[8], evDone[8];
    cudaStream_t side; cudaEvent_t evFork, evJoin;   // the blur runs beside the quadtree (it needs only the pyramid)
    bool pending;                        // an orbb200_extract_host_async call has not been waited for yet
    std::map<int, cudaGraphExec_t> graphs;   // host path, small batches: the kernel sequence of one call as a CUDA graph, by batch size
    std::map<int, int> graphLaunches;                // kernels inside each of them
    std::vector<void*> allocs;
};

static int round_half_even(float v) { return (int)lrintf(v); }

static short coef11(float c)
{
    int v = round_half_even(c * 2048.f);
    return (short)std::min(std::max(v, -32768), 32767);
}

// cv::resize INTER_LINEAR coefficient tables for sw x sh -> dw x dh (see oracle/orb_oracle.c)
static void build_resize_tables(int sw, int sh, int dw, int dh, std::vector<short4>& xt, std::vector<short4>& yt)
{
    const double scale_x = 1. / ((double)dw / sw), scale_y = 1. / ((double)dh / sh);
    xt.resize(dw + 4); yt.resize(dh);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xt[dx] = make_short4((short)sx, coef11(1.f - fx), coef11(fx), (short)std::min(sx + 1, sw - 1));
    }
    for (int i = 0; i < 4; i++) xt[dw + i] = xt[dw - 2 - i];    // columns dw..dw+3 mirror dw-2..dw-5 (REFLECT_101)
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        const int y0 = std::min(std::max(sy, 0), sh - 1), y1 = std::min(std::max(sy + 1, 0), sh - 1);
        yt[dy] = make_short4((short)y0, (short)y1, coef11(1.f - fy), coef11(fy));
    }
}

// 3-D tensor map {column, row, frame} over one pyramid level for k_fast2's window copies (uint8, no swizzle, zero fill).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int encode_level_map(CUtensorMap* map, const void* base, int w, int h, int frames, size_t pitch, size_t frameStride, int boxW, int boxH, int boxFrames = 1)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        ORB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return ORBB200_ECUDA; }
        fn = (EncodeTiledFn)p;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frameStride};
    const cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, (cuuint32_t)boxFrames};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for a %d x %d x %d level, pitch %zu, frame stride %zu", (int)r, w, h, frames, pitch, frameStride); return ORBB200_ECUDA; }
    return ORBB200_OK;
}

template <typename T>
static int dev_alloc(orbb200_extractor* h, T** p, size_t count)
{
    void* q = nullptr;
    ORB_CUDA(cudaMalloc(&q, count * sizeof(T) + 256));      // slack: kernels read whole aligned words near row ends
    h->allocs.push_back(q);
    *p = (T*)q;
    return ORBB200_OK;
}

extern "C" const char* orbb200_last_error(void) { return g_err; }
extern "C" const char* orbb200_version(void) { return "orb_b200 0.1 (sm_100a)"; }
extern "C" int orbb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int orbb200_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST,
                                        int width, int height, int max_batch, int device, int blur_taps,
                                        orbb200_extractor** out)
{
    if (!out) { set_error("out is NULL"); return ORBB200_EINVAL; }
    *out = nullptr;
    if (nlevels < 1 || nlevels > MAXL || nfeatures < 0 || max_batch < 1 || !(scaleFactor > 1.0f) || scaleFactor > 2.0f ||
        width < 1 || height < 1 || width > MAX_DIM || height > MAX_DIM || iniThFAST < 0 || minThFAST < 0 ||
        iniThFAST > 255 || minThFAST > 255) {
        set_error("invalid extractor parameters");
        return ORBB200_EINVAL;
    }
    int ndev = orbb200_device_count();
    if (device < 0 || device >= ndev) { set_error("CUDA device %d not available (%d visible)", device, ndev); return ORBB200_ENODEVICE; }
    ORB_CUDA(cudaSetDevice(device));

    orbb200_extractor* h = new orbb200_extractor();
    h->nfeatures = nfeatures; h->nlevels = nlevels; h->iniTh = iniThFAST; h->minTh = minThFAST;
    h->width = width; h->height = height; h->maxBatch = max_batch; h->device = device; h->blurTaps = blur_taps ? 1 : 0;
    h->pinned = nullptr; h->pinnedBytes = 0; h->lastLaunches = 0; h->lastBatch = 0; h->lastIn = nullptr;
    h->profiling = false;
    h->fastSmallCells = 0;
    h->chunkOverride = 0;
    if (const char* e = getenv("ORBB200_CHUNKS")) h->chunkOverride = std::max(0, std::min(8, atoi(e)));
    for (int i = 0; i < 6; i++) h->ev[i] = nullptr;
    h->copyIn = h->copyOut = nullptr;
    for (int i = 0; i < 8; i++) h->evIn[i] = h->evDone[i] = nullptr;
    // ---- constructor tables (S/ORBextractor.cc:421-455); scaleFactor is held in a double (I/ORBextractor.h:98)
    h->scaleFactorD = (double)scaleFactor;
    h->scale[0] = 1.0f; h->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        h->scale[i] = (float)((double)h->scale[i - 1] * h->scaleFactorD);
        h->sigma2[i] = h->scale[i] * h->scale[i];
    }
    for (int i = 0; i < nlevels; i++) { h->invScale[i] = 1.0f / h->scale[i]; h->invSigma2[i] = 1.0f / h->sigma2[i]; }
    {
        const float factor = (float)(1.0 / h->scaleFactorD);
        float desired = (float)nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
        int sum = 0;
        for (int l = 0; l < nlevels - 1; l++) {
            h->perLevel[l] = round_half_even(desired);
            sum += h->perLevel[l];
            desired *= factor;
        }
        h->perLevel[nlevels - 1] = std::max(nfeatures - sum, 0);
    }

    // ---- per-level geometry
    ExtractParams& P = h->P;
    memset(&P, 0, sizeof(P));
    P.nlevels = nlevels; P.iniTh = iniThFAST; P.minTh = minThFAST; P.blurVariant = h->blurTaps;
    std::vector<short4> tabs;
    std::vector<int> rzXs;
    std::vector<int4> ytab3;
    long long pyrOff = 0, blurOff = 0;
    int cells = 0, candOff = 0, kpOff = 0, tiles = 0, maxWCell = 0, maxHCell = 0, maxKpCap = 0;
    for (int l = 0; l < nlevels; l++) {
        LevelGeo& g = P.lv[l];
        g.w = round_half_even((float)width * h->invScale[l]);      // :1145
        g.h = round_half_even((float)height * h->invScale[l]);
        g.pitch = (int)align_up(g.w + 4, 16);       // >= 4 spare columns: k_resize stores the REFLECT_101 continuation there
        g.maxBX = g.w - EDGE + 3; g.maxBY = g.h - EDGE + 3;
        const float fw = (float)(g.maxBX - BORDER), fh = (float)(g.maxBY - BORDER);
        if (fw < 30.f || fh < 30.f) {
            set_error("level %d is %dx%d: narrower than one 30-px FAST cell plus borders; the reference divides by zero here", l, g.w, g.h);
            delete h; return ORBB200_EGEOMETRY;
        }
        g.nCols = (int)(fw / 30.f); g.nRows = (int)(fh / 30.f);                    // :797-798
        g.wCell = (int)ceilf(fw / g.nCols); g.hCell = (int)ceilf(fh / g.nRows);    // :799-800
        g.cellStart = cells; cells += g.nCols * g.nRows;
        maxWCell = std::max(maxWCell, g.wCell); maxHCell = std::max(maxHCell, g.hCell);
        g.N = h->perLevel[l];
        g.nIni = (int)roundf((float)(g.maxBX - BORDER) / (g.maxBY - BORDER));      // :556
        if (g.nIni < 1) {
            set_error("level %d aspect ratio %d:%d rounds to zero quadtree roots; the reference divides by zero here", l, g.w, g.h);
            delete h; return ORBB200_EGEOMETRY;
        }
        g.hX = (float)(g.maxBX - BORDER) / g.nIni;                                  // :558
        // 3x3 NMS with a strict ">" keeps at most one corner per 2x2 block of a cell's detection area, so the sum over
        // the cells of ceil(dw / 2) * ceil(dh / 2) bounds the level's candidates (STATUS_CAND_OVERFLOW cannot be reached
        // unless the capacity is shrunk on purpose, orbb200_extractor_debug_set_capacity)
        g.candCap = 0;
        for (int ci = 0; ci < g.nRows; ci++)
            for (int cj = 0; cj < g.nCols; cj++) {
                const int iniY = BORDER + ci * g.hCell, iniX = BORDER + cj * g.wCell;
                if (iniY >= g.maxBY - 3 || iniX >= g.maxBX - 6) continue;
                const int dw = std::min(iniX + g.wCell + 6, g.maxBX) - iniX - 6, dh = std::min(iniY + g.hCell + 6, g.maxBY) - iniY - 6;
                if (dw > 0 && dh > 0) g.candCap += ((dw + 1) / 2) * ((dh + 1) / 2);
            }
        g.candCap = std::max(g.candCap, 64);
        g.candOff = candOff; candOff += (int)align_up(g.candCap, 64);
        g.kpCap = std::max(g.N + 3, 4 * g.nIni);
        g.kpOff = kpOff; kpOff += g.kpCap;
        maxKpCap = std::max(maxKpCap, g.kpCap + g.nIni);
        g.scale = h->scale[l];
        g.kpSize = (float)(int)(31 * h->scale[l]);                                   // :855
        g.blurOff = blurOff; blurOff += (long long)align_up((size_t)g.pitch * g.h, 256);
        if (l > 0) {
            g.pyrOff = pyrOff; pyrOff += (long long)align_up((size_t)g.pitch * g.h, 256);
            std::vector<short4> xt, yt;
            build_resize_tables(P.lv[l - 1].w, P.lv[l - 1].h, g.w, g.h, xt, yt);
            for (int x0 = 0; x0 < g.w + 4; x0 += 4) {      // k_resize cuts 4 columns out of one 8-byte window
                int lo = 1 << 30, hi = 0;
                for (int j = 0; j < 4; j++) { const short4 e = xt[std::min(x0 + j, g.w + 3)]; lo = std::min(lo, (int)e.x); hi = std::max(hi, (int)e.w); }
                if (hi - lo > 7) { set_error("scale factor too large for the resize kernel (column span %d)", hi - lo); delete h; return ORBB200_EGEOMETRY; }
            }
            g.xtabOff = (int)tabs.size(); tabs.insert(tabs.end(), xt.begin(), xt.end());
            g.ytabOff = (int)tabs.size(); tabs.insert(tabs.end(), yt.begin(), yt.end());
            // k_resize3: per 64-column block the 16-aligned first source column and the box that holds any block's source window, and the row table {lower source row, b0 << 16, b1 << 16, both taps on one row}
            P.rz3XsOff[l] = (int)rzXs.size();
            int boxW3 = 0, boxH3 = 0;
            for (int bx = 0; bx * 64 < g.w + 4; bx++) {
                int lo = 1 << 30, hi = 0;
                for (int x0 = bx * 64; x0 < std::min(bx * 64 + 64, g.w + 4); x0 += 4) {
                    int bmin = 1 << 30;
                    for (int j = 0; j < 4; j++) bmin = std::min(bmin, (int)xt[std::min(x0 + j, g.w + 3)].x);
                    lo = std::min(lo, bmin & ~3); hi = std::max(hi, (bmin & ~3) + 12);
                }
                rzXs.push_back(lo & ~15);
                boxW3 = std::max(boxW3, hi - (lo & ~15));
            }
            bool monotone = true;
            for (int y = 0; y < g.h; y++)
                if (yt[y].y != yt[y].x + 1 || (y > 0 && yt[y].y < yt[y - 1].y)) monotone = false;
            P.ytab3Off[l] = (int)ytab3.size();
            for (int y0 = 0; y0 < g.h; y0 += RZ3_ROWS) {
                for (int y = y0; y < y0 + RZ3_ROWS; y++) {
                    const short4 e = yt[std::min(y, g.h - 1)];
                    if (y < g.h) ytab3.push_back(make_int4((int)e.y, (int)((uint32_t)(uint16_t)e.z << 16), (int)((uint32_t)(uint16_t)e.w << 16), 0));
                    else ytab3.push_back(make_int4(-1, 0, 0, 0));
                }
                ytab3.push_back(make_int4(-1, 0, 0, 0));
                boxH3 = std::max(boxH3, (int)yt[std::min(y0 + RZ3_ROWS, g.h) - 1].y - ((int)yt[y0].y - 1) + 1);
            }
            boxW3 = (int)align_up(boxW3, 16);
            P.rz3BoxW[l] = (monotone && boxW3 <= 256 && boxH3 <= 256) ? boxW3 : 0;
            P.rz3BoxH[l] = boxH3;
        }
        g.blurTilesX = (g.w + 127) / 128;
        g.blurTileStart = tiles; tiles += g.blurTilesX * ((g.h + BL_ROWS - 1) / BL_ROWS);
    }
    for (int l = 0; l < nlevels; l++) h->realCandCap[l] = P.lv[l].candCap;
    h->totalCells = cells; h->totalBlurTiles = tiles; h->maxKp = kpOff;
    P.totalBlurTiles = tiles;
    P.pyrFrameBytes = std::max<long long>(pyrOff, 256); P.blurFrameBytes = blurOff;
    P.candFrameCap = candOff; P.kpFrameCap = kpOff; P.outCap = kpOff;

    // k_fast: pick the tile instantiation that holds the largest cell (+6 halo, +6 alignment slack)
    if (maxWCell > 57 || maxHCell > 57) { set_error("FAST cell larger than 57 px"); delete h; return ORBB200_EGEOMETRY; }
    P.fastLarge = FastGeo<26, 42>::fits(maxWCell, maxHCell) ? 0 : 1;
    if (P.fastLarge && !FastGeo<38, 64>::fits(maxWCell, maxHCell)) { set_error("FAST cell does not fit the staging tile"); delete h; return ORBB200_EGEOMETRY; }
    P.totalCells = cells;
    h->fastVariant = 2;
    if (const char* e = getenv("ORBB200_FAST_VARIANT")) h->fastVariant = atoi(e) == 1 ? 1 : 2;
    if (h->fastVariant == 2) {
        P.fastLarge = FastGeo2<26, 42>::fits(maxWCell, maxHCell) ? 0 : 1;
        if (P.fastLarge && !FastGeo2<38, 64>::fits(maxWCell, maxHCell)) { set_error("FAST cell does not fit the staging tile"); delete h; return ORBB200_EGEOMETRY; }
        h->fastSmem = P.fastLarge ? FastGeo2<38, 64>::SMEM_BYTES : FastGeo2<26, 42>::SMEM_BYTES;
    } else
    h->fastSmem = (size_t)FAST_WARPS * (P.fastLarge ? FastGeo<38, 64>::WARP_BYTES : FastGeo<26, 42>::WARP_BYTES);
    // k_fast2's cell table: window origin and size, level, offset of the cell inside the level (:807-822, :840-841)
    // (levels whose cells fit the 22-word tile come first: they run in their own launch, see launch_kernels)
    bool smallTile = h->fastVariant == 2 && !P.fastLarge;
    if (const char* e = getenv("ORBB200_FAST_TILE")) if (atoi(e) == 26) smallTile = false;      // A/B switch: the 26-word tile for every cell
    std::vector<int4> cellTab;
    for (int pass = 0; pass < 2; pass++) {
    if (pass == 1) h->fastSmallCells = (int)cellTab.size();
    for (int l = 0; l < nlevels; l++) {
        const LevelGeo& g = P.lv[l];
        if ((smallTile && FastGeo2<22, 42>::fits(g.wCell, g.hCell)) != (pass == 0)) continue;
        for (int ci = 0; ci < g.nRows; ci++)
            for (int cj = 0; cj < g.nCols; cj++) {
                const int iniY = BORDER + ci * g.hCell, iniX = BORDER + cj * g.wCell;
                if (iniY >= g.maxBY - 3 || iniX >= g.maxBX - 6) continue;                    // :810, :819
                const int ww = std::min(iniX + g.wCell + 6, g.maxBX) - iniX, wh = std::min(iniY + g.hCell + 6, g.maxBY) - iniY;
                if (ww < 7 || wh < 7) continue;                                               // cv::FAST finds nothing in < 7 px
                cellTab.push_back(make_int4(iniX | (iniY << 16), ww | (wh << 8) | (l << 16), (cj * g.wCell) | ((ci * g.hCell) << 16), 0));
            }
    }
    }
    P.nCells = (int)cellTab.size();
    // k_quadtree shared memory: 88 bytes per node slot + 6 bytes per candidate held on chip
    P.qtNC = (int)align_up(maxKpCap + 8, 8);
    const size_t perNode = 8 + 8 + 8 + 4 * 7 + 16 + 16;
    const size_t budget = 200 * 1024;
    if (perNode * P.qtNC + 6 * 1024 > budget) { set_error("nfeatures too large for the on-chip quadtree (%d node slots)", P.qtNC); delete h; return ORBB200_EINVAL; }
    P.qtPC = (int)std::min<size_t>(QT_POINTS_ON_CHIP, (budget - perNode * P.qtNC) / 6) & ~7;
    h->qtSmem = perNode * P.qtNC + 6 * (size_t)P.qtPC;

    // ---- device memory
    // staging slab for host frames / non-canonical device frames: tight rows when the width allows 16-byte
    // chunks (host copies are then linear, which is what PCIe DMA likes), padded rows otherwise
    h->inPitch = (width % 16 == 0) ? (size_t)width : align_up(width + 4, 16);
    int rc;
#define TRY(x) do { rc = (x); if (rc != ORBB200_OK) { orbb200_extractor_destroy(h); return rc; } } while (0)
    short4* dTabs = nullptr;
    TRY(dev_alloc(h, &h->dIn, h->inPitch * height * (size_t)max_batch));
    TRY(dev_alloc(h, &P.pyr, (size_t)P.pyrFrameBytes * max_batch));
    TRY(dev_alloc(h, &P.blur, (size_t)P.blurFrameBytes * max_batch));
    TRY(dev_alloc(h, &dTabs, tabs.size() + 1));
    TRY(dev_alloc(h, &P.cand, (size_t)P.candFrameCap * max_batch));
    TRY(dev_alloc(h, &P.qtScratch, (size_t)P.candFrameCap * max_batch));
    TRY(dev_alloc(h, &P.candCount, (size_t)nlevels * max_batch));
    TRY(dev_alloc(h, &P.lkp, (size_t)P.kpFrameCap * max_batch));
    TRY(dev_alloc(h, &P.lkpCount, (size_t)nlevels * max_batch));
    TRY(dev_alloc(h, &h->dOutKp, (size_t)P.outCap * max_batch));
    TRY(dev_alloc(h, &h->dOutDesc, (size_t)P.outCap * max_batch * 32));
    TRY(dev_alloc(h, &h->dOutCount, (size_t)max_batch));
    TRY(dev_alloc(h, &P.status, 1));
    int4* dCells = nullptr;
    TRY(dev_alloc(h, &dCells, cellTab.size() + 1));
    P.cells = dCells;
    int* dRzXs = nullptr;
    TRY(dev_alloc(h, &dRzXs, rzXs.size() + 1));
    P.rzXs = dRzXs;
    int4* dYtab3 = nullptr;
    TRY(dev_alloc(h, &dYtab3, ytab3.size() + 1));
    P.ytab3 = dYtab3;
    h->resizeVariant = 3;
    if (const char* ev = getenv("ORBB200_RESIZE_VARIANT")) h->resizeVariant = atoi(ev) == 1 ? 1 : 3;
    for (int l = 2; l < nlevels && h->resizeVariant == 3; l++)      // source = level l - 1 in the handle's pyramid slab
        if (P.rz3BoxW[l]) TRY(encode_level_map(&h->resizeMaps.m[l], P.pyr + P.lv[l - 1].pyrOff, P.lv[l - 1].w + 4, P.lv[l - 1].h, max_batch, P.lv[l - 1].pitch,
                                              (size_t)P.pyrFrameBytes, P.rz3BoxW[l], P.rz3BoxH[l], 2));
    std::vector<uint32_t> blurTab;
    for (int l = 0; l < nlevels; l++)
        for (int ty = 0; ty < (P.lv[l].h + BL_ROWS - 1) / BL_ROWS; ty++)
            for (int tx = 0; tx < P.lv[l].blurTilesX; tx++) blurTab.push_back(((uint32_t)l << 24) | ((uint32_t)ty << 12) | (uint32_t)tx);
    uint32_t* dBlurTiles = nullptr;
    TRY(dev_alloc(h, &dBlurTiles, blurTab.size() + 1));
    P.blurTiles = dBlurTiles;
    h->descVariant = 3;
    if (const char* ev = getenv("ORBB200_DESCRIBE_VARIANT")) h->descVariant = atoi(ev) == 1 ? 1 : 3;
    h->maxLevelKpCap = 0;
    for (int l = 0; l < nlevels; l++) h->maxLevelKpCap = std::max(h->maxLevelKpCap, P.lv[l].kpCap);
    for (int l = 0; l < nlevels && h->descVariant == 3; l++) {
        if (l > 0) TRY(encode_level_map(&h->descMaps.u[l], P.pyr + P.lv[l].pyrOff, P.lv[l].w, P.lv[l].h, max_batch, P.lv[l].pitch, (size_t)P.pyrFrameBytes, DESC_UW, DESC_UROWS));
        TRY(encode_level_map(&h->descMaps.b[l], P.blur + P.lv[l].blurOff, P.lv[l].w, P.lv[l].h, max_batch, P.lv[l].pitch, (size_t)P.blurFrameBytes, DESC_BW, DESC_BROWS));
    }
    for (int l = 1; l < nlevels; l++)
        TRY(encode_level_map(&h->blurMaps.m[l], P.pyr + P.lv[l].pyrOff, P.lv[l].pitch, P.lv[l].h, max_batch, P.lv[l].pitch, (size_t)P.pyrFrameBytes, BL_ROW_BYTES, BL_IN_ROWS));
    for (int l = 1; l < nlevels && h->fastVariant == 2; l++)      // levels >= 1 live in the handle's pyramid slab: maps made once
        TRY(encode_level_map(&h->fastMaps.m[l], P.pyr + P.lv[l].pyrOff, P.lv[l].w, P.lv[l].h, max_batch, P.lv[l].pitch, (size_t)P.pyrFrameBytes,
                             P.fastLarge ? FastGeo2<38, 64>::BW : FastGeo2<26, 42>::BW, P.fastLarge ? 64 : 42));
#undef TRY
    P.tabs = dTabs;
    cudaError_t e = cudaSuccess;
    if (!cellTab.empty()) e = cudaMemcpy(dCells, cellTab.data(), cellTab.size() * sizeof(int4), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !blurTab.empty()) e = cudaMemcpy(dBlurTiles, blurTab.data(), blurTab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !rzXs.empty()) e = cudaMemcpy(dRzXs, rzXs.data(), rzXs.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !ytab3.empty()) e = cudaMemcpy(dYtab3, ytab3.data(), ytab3.size() * sizeof(int4), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("extractor_create: %s", cudaGetErrorString(e)); orbb200_extractor_destroy(h); return ORBB200_ECUDA; }
    if (!tabs.empty()) e = cudaMemcpy(dTabs, tabs.data(), tabs.size() * sizeof(short4), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(P.status, 0, sizeof(int));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copyIn, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copyOut, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->evFork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->evJoin, cudaEventDisableTiming);
    for (int i = 0; i < 8 && e == cudaSuccess; i++) {
        e = cudaEventCreateWithFlags(&h->evIn[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->evDone[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess && h->fastVariant == 1) e = ensure_dynamic_smem(P.fastLarge ? (const void*)k_fast<38, 64> : (const void*)k_fast<26, 42>, device, h->fastSmem);
    if (e == cudaSuccess && h->fastVariant == 2) e = ensure_dynamic_smem(P.fastLarge ? (const void*)k_fast2<38, 64> : (const void*)k_fast2<26, 42>, device, h->fastSmem);
    if (e == cudaSuccess && h->descVariant == 3) e = ensure_dynamic_smem((const void*)k_describe3, device, D3_CTA_BYTES);
    for (int l = 1; l < nlevels && e == cudaSuccess && h->resizeVariant == 3; l++)
        if (P.rz3BoxW[l]) e = ensure_dynamic_smem((const void*)k_resize3, device, resize3_smem(P, l));
    if (e == cudaSuccess) e = ensure_dynamic_smem((const void*)k_quadtree, device, h->qtSmem);
    if (e == cudaSuccess) e = ensure_dynamic_smem(P.blurVariant ? (const void*)k_blur<true> : (const void*)k_blur<false>, device, BL_WARPS * BL_WARP_BYTES);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->numSMs, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) {
        set_error("extractor_create: %s", cudaGetErrorString(e));
        orbb200_extractor_destroy(h);
        return ORBB200_ECUDA;
    }
    *out = h;
    return ORBB200_OK;
}

extern "C" void orbb200_extractor_destroy(orbb200_extractor* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    for (std::map<int, cudaGraphExec_t>::iterator it = h->graphs.begin(); it != h->graphs.end(); ++it) cudaGraphExecDestroy(it->second);
    for (int i = 0; i < 6; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i < 8; i++) { if (h->evIn[i]) cudaEventDestroy(h->evIn[i]); if (h->evDone[i]) cudaEventDestroy(h->evDone[i]); }
    if (h->copyIn) cudaStreamDestroy(h->copyIn);
    if (h->copyOut) cudaStreamDestroy(h->copyOut);
    if (h->side) { cudaStreamSynchronize(h->side); cudaStreamDestroy(h->side); }
    if (h->evFork) cudaEventDestroy(h->evFork);
    if (h->evJoin) cudaEventDestroy(h->evJoin);
    for (void* p : h->allocs) cudaFree(p);
    if (h->pinned) cudaFreeHost(h->pinned);
    delete h;
}

extern "C" int orbb200_extractor_tables(const orbb200_extractor* h, float* scale, float* inv_scale, float* sigma2,
                                        float* inv_sigma2, int* features_per_level, int* umax16)
{
    if (!h) { set_error("null handle"); return ORBB200_EINVAL; }
    static const int umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    for (int i = 0; i < h->nlevels; i++) {
        if (scale) scale[i] = h->scale[i];
        if (inv_scale) inv_scale[i] = h->invScale[i];
        if (sigma2) sigma2[i] = h->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = h->invSigma2[i];
        if (features_per_level) features_per_level[i] = h->perLevel[i];
    }
    if (umax16) memcpy(umax16, umax, sizeof(umax));
    return ORBB200_OK;
}
extern "C" int orbb200_extractor_levels(const orbb200_extractor* h) { return h ? h->nlevels : ORBB200_EINVAL; }
extern "C" int orbb200_extractor_max_keypoints(const orbb200_extractor* h) { return h ? h->maxKp : ORBB200_EINVAL; }
extern "C" int orbb200_extractor_level_size(const orbb200_extractor* h, int level, int* w, int* hgt)
{
    if (!h || level < 0 || level >= h->nlevels) { set_error("bad level"); return ORBB200_EINVAL; }
    if (w) *w = h->P.lv[level].w;
    if (hgt) *hgt = h->P.lv[level].h;
    return ORBB200_OK;
}
extern "C" void* orbb200_extractor_stream(orbb200_extractor* h) { return h ? (void*)h->stream : nullptr; }
extern "C" int orbb200_extractor_last_launches(const orbb200_extractor* h) { return h ? h->lastLaunches : 0; }
extern "C" int orbb200_extractor_device(const orbb200_extractor* h) { return h ? h->device : ORBB200_EINVAL; }
extern "C" int orbb200_extractor_debug_set_capacity(orbb200_extractor* h, int candidates_per_level)
{
    if (!h) { set_error("null handle"); return ORBB200_EINVAL; }
    for (int l = 0; l < h->nlevels; l++)
        h->P.lv[l].candCap = candidates_per_level > 0 ? std::min(h->realCandCap[l], candidates_per_level) : h->realCandCap[l];
    for (std::map<int, cudaGraphExec_t>::iterator it = h->graphs.begin(); it != h->graphs.end(); ++it) cudaGraphExecDestroy(it->second);
    h->graphs.clear();          // captured launches carry the old geometry
    return ORBB200_OK;
}

extern "C" int orbb200_extractor_set_profiling(orbb200_extractor* h, int on)
{
    if (!h) { set_error("null handle"); return ORBB200_EINVAL; }
    ORB_CUDA(cudaSetDevice(h->device));
    if (on) for (int i = 0; i < 6; i++) if (!h->ev[i]) ORB_CUDA(cudaEventCreate(&h->ev[i]));
    h->profiling = on != 0;
    return ORBB200_OK;
}

extern "C" int orbb200_extractor_stage_ms(orbb200_extractor* h, float* ms5)
{
    if (!h || !ms5 || !h->profiling) { set_error("profiling is off"); return ORBB200_EINVAL; }
    ORB_CUDA(cudaSetDevice(h->device));
    ORB_CUDA(cudaEventSynchronize(h->ev[5]));
    for (int i = 0; i < 5; i++) ORB_CUDA(cudaEventElapsedTime(&ms5[i], h->ev[i], h->ev[i + 1]));
    return ORBB200_OK;
}

// The kernel sequence of one extraction call on stream st (also what a CUDA graph of the call contains).
static int launch_kernels(orbb200_extractor* h, const ExtractParams& P, int batch, cudaStream_t st, int& launches)
{
    if (P.inPadded) {       // level 0 sits in the padded staging slab: write the REFLECT_101 continuation columns the blur reads
        k_pad_level0<<<dim3((h->height + 127) / 128, batch), 128, 0, st>>>(const_cast<uint8_t*>(P.in), P.inFrameStride, P.inPitch, h->width, h->height);
        ORB_CHECK_LAUNCH("k_pad_level0");
    }
    ORB_CUDA(cudaMemsetAsync(P.candCount, 0, sizeof(int) * P.nlevels * batch, st));
#define STAGE_MARK(i) do { if (h->profiling) ORB_CUDA(cudaEventRecord(h->ev[i], st)); } while (0)
    STAGE_MARK(0);
    if (h->resizeVariant == 3 && P.nlevels > 1 && P.rz3BoxW[1]) {
        int rc = encode_level_map(&h->resizeMaps.m[1], P.in, h->width, h->height, batch, (size_t)P.inPitch, (size_t)P.inFrameStride, P.rz3BoxW[1], P.rz3BoxH[1], 2);
        if (rc != ORBB200_OK) return rc;
    }
    for (int l = 1; l < P.nlevels; l++) {
        const LevelGeo& g = P.lv[l];
        dim3 grid((g.w + 4 + 127) / 128, (g.h + RZ_ROWS * RZ_WARPS - 1) / (RZ_ROWS * RZ_WARPS), batch);
        if (h->resizeVariant == 3 && P.rz3BoxW[l])
            launch_pdl(k_resize3, dim3((g.w + 4 + 63) / 64, (g.h + RZ3_ROWS * RZ3_WARPS - 1) / (RZ3_ROWS * RZ3_WARPS), (batch + 1) / 2), RZ3_WARPS * 32, resize3_smem(P, l), st, P, h->resizeMaps, l);
        else k_resize<<<grid, RZ_WARPS * 32, 0, st>>>(P, l);
        ORB_CHECK_LAUNCH("k_resize"); launches++;
    }
    STAGE_MARK(1);
    if (h->fastVariant == 2) {
        // level 0 may be the caller's own buffer: its map is encoded per call (a host-side table fill, no driver round trip)
        int rc = encode_level_map(&h->fastMaps.m[0], P.in, h->width, h->height, batch, (size_t)P.inPitch, (size_t)P.inFrameStride,
                                  P.fastLarge ? FastGeo2<38, 64>::BW : FastGeo2<26, 42>::BW, P.fastLarge ? 64 : 42);
        if (rc != ORBB200_OK) return rc;
        const int nSmall = h->fastSmallCells;
        if (P.fastLarge) k_fast2<38, 64><<<dim3(P.nCells, batch), 32, h->fastSmem, st>>>(P, h->fastMaps);
        else {
            // the few cells wider than 33 px first, in the 26-word tile; the others follow in the 22-word tile as a dependent launch
            // that waits for nothing (the two grids touch different levels): its blocks fill the SMs as the first grid drains
            if (nSmall < P.nCells) {
                ExtractParams Pw = P;
                Pw.cells += nSmall;
                k_fast2<26, 42><<<dim3(P.nCells - nSmall, batch), 32, FastGeo2<26, 42>::SMEM_BYTES, st>>>(Pw, h->fastMaps);
                ORB_CHECK_LAUNCH("k_fast"); launches++;
            }
            if (nSmall > 0 && nSmall < P.nCells) {
                const cudaError_t le = launch_pdl(k_fast2<22, 42>, dim3(nSmall, batch), 32, FastGeo2<22, 42>::SMEM_BYTES, st, P, h->fastMaps);
                if (le != cudaSuccess) { set_error("launch of k_fast failed: %s", cudaGetErrorString(le)); return ORBB200_ECUDA; }
            } else if (nSmall > 0)      // (alone it must wait for the pyramid like any kernel: a plain launch)
                k_fast2<22, 42><<<dim3(nSmall, batch), 32, FastGeo2<22, 42>::SMEM_BYTES, st>>>(P, h->fastMaps);
        }
    } else {
        const dim3 grid((h->totalCells + FAST_WARPS - 1) / FAST_WARPS, batch);
        if (P.fastLarge) k_fast<38, 64><<<grid, FAST_WARPS * 32, h->fastSmem, st>>>(P);
        else k_fast<26, 42><<<grid, FAST_WARPS * 32, h->fastSmem, st>>>(P);
    }
    ORB_CHECK_LAUNCH("k_fast"); launches++;
    STAGE_MARK(2);
    // The blur needs only the pyramid and the quadtree only FAST's candidates.  In a small batch neither fills the GPU (the
    // quadtree is one CTA per tree and a chain of barriers), so there the blur's warps run beside the quadtree on a side
    // stream: fork after FAST, join before the descriptors (one frame per call: 0.168 -> 0.156 ms).  A full batch keeps
    // the two in sequence -- measured at 256 frames, side by side they take 0.406 ms instead of 0.368 ms.
    // With stage events on, a forked "blur" stage is what is left of the blur once the quadtree has finished.
    const bool beside = batch <= SIDE_MAX_BATCH;
    cudaStream_t bs = beside ? h->side : st;
    if (beside) {
        ORB_CUDA(cudaEventRecord(h->evFork, st));
        ORB_CUDA(cudaStreamWaitEvent(h->side, h->evFork, 0));
    } else {
        k_quadtree<<<dim3(P.nlevels, batch), QT_THREADS, h->qtSmem, st>>>(P);
        ORB_CHECK_LAUNCH("k_quadtree"); launches++;
        STAGE_MARK(3);
    }
    {   // persistent warps: a few CTAs per SM walk the (frame, level, tile) list
        const long long tiles = (long long)h->totalBlurTiles * batch;
        const int ctas = (int)std::min<long long>((tiles + BL_WARPS - 1) / BL_WARPS, (long long)h->numSMs * BL_CTAS_PER_SM);
        int rc = encode_level_map(&h->blurMaps.m[0], P.in, P.inRowBytes, h->height, batch, (size_t)P.inPitch, (size_t)P.inFrameStride, BL_ROW_BYTES, BL_IN_ROWS);
        if (rc != ORBB200_OK) return rc;
        if (P.blurVariant) k_blur<true><<<ctas, BL_WARPS * 32, BL_WARPS * BL_WARP_BYTES, bs>>>(P, h->blurMaps);
        else k_blur<false><<<ctas, BL_WARPS * 32, BL_WARPS * BL_WARP_BYTES, bs>>>(P, h->blurMaps);
    }
    ORB_CHECK_LAUNCH("k_blur"); launches++;
    if (beside) {
        ORB_CUDA(cudaEventRecord(h->evJoin, h->side));
        k_quadtree<<<dim3(P.nlevels, batch), QT_THREADS, h->qtSmem, st>>>(P);
        ORB_CHECK_LAUNCH("k_quadtree"); launches++;
        STAGE_MARK(3);
        ORB_CUDA(cudaStreamWaitEvent(st, h->evJoin, 0));
    }
    STAGE_MARK(4);
    if (h->descVariant == 3) {
        int rc = encode_level_map(&h->descMaps.u[0], P.in, h->width, h->height, batch, (size_t)P.inPitch, (size_t)P.inFrameStride, DESC_UW, DESC_UROWS);
        if (rc != ORBB200_OK) return rc;
        // persistent warps over groups of 8 output rows; calls of a few frames (latency counts): one row per group
        ExtractParams Pd = P;
        Pd.descChunk = batch <= SIDE_MAX_BATCH ? 1 : 8;
        const int items = batch * ((P.outCap + Pd.descChunk - 1) / Pd.descChunk);
        k_describe3<<<std::max(1, std::min(h->numSMs * DESC_CTAS_PER_SM, (items + DESC_WARPS - 1) / DESC_WARPS)), DESC_WARPS * 32, D3_CTA_BYTES, st>>>(Pd, h->descMaps);
    } else
    k_describe<<<dim3((P.kpFrameCap + DESC_WARPS - 1) / DESC_WARPS, batch), DESC_WARPS * 32, 0, st>>>(P);
    ORB_CHECK_LAUNCH("k_describe"); launches++;
    STAGE_MARK(5);
#undef STAGE_MARK
    return ORBB200_OK;
}

// enqueue the whole pipeline for `batch` frames whose level 0 is at d_images
// `first` = index of the scratch slabs (pyramid, candidates, ...) the batch's frame 0 uses, so that several
// chunks of one host call can be in flight in disjoint parts of the handle's buffers.
static int enqueue(orbb200_extractor* h, const uint8_t* d_images, int batch, size_t stride, size_t frame_stride,
                   orbb200_keypoint* d_kp, uint8_t* d_desc, int32_t* d_counts, int cap, int first = 0)
{
    ExtractParams P = h->P;
    cudaStream_t st = h->stream;
    P.pyr += (size_t)first * P.pyrFrameBytes; P.blur += (size_t)first * P.blurFrameBytes;
    P.cand += (size_t)first * P.candFrameCap; P.qtScratch += (size_t)first * P.candFrameCap;
    P.candCount += (size_t)first * P.nlevels; P.lkp += (size_t)first * P.kpFrameCap; P.lkpCount += (size_t)first * P.nlevels;
    const size_t inFrameBytes = h->inPitch * (size_t)h->height;
    const bool staged = d_images >= h->dIn && d_images < h->dIn + inFrameBytes * h->maxBatch;
    // The kernels read level 0 as aligned words and 16-byte TMA bulk rows.  A caller buffer qualifies when base,
    // strides and width are multiples of 16 (then no chunk straddles the end of a row); anything else is
    // first copied into the handle's padded staging slab.
    if (!staged) {
        const bool canonical = ((reinterpret_cast<uintptr_t>(d_images) | stride | frame_stride | (size_t)h->width) & 15) == 0;
        if (canonical) {
            P.inRowBytes = h->width;
        } else {
            for (int f = 0; f < batch; f++)
                ORB_CUDA(cudaMemcpy2DAsync(h->dIn + (size_t)(first + f) * inFrameBytes, h->inPitch, d_images + f * frame_stride,
                                           stride, h->width, h->height, cudaMemcpyDeviceToDevice, st));
            d_images = h->dIn + (size_t)first * inFrameBytes; stride = h->inPitch; frame_stride = inFrameBytes;
            P.inRowBytes = (int)h->inPitch;
        }
    } else {
        P.inRowBytes = (int)h->inPitch;
    }
    P.batch = batch; P.in = d_images; P.inPitch = (int)stride; P.inFrameStride = (long long)frame_stride;
    P.frameBase = first;
    P.inPadded = (d_images >= h->dIn && d_images < h->dIn + inFrameBytes * h->maxBatch && h->inPitch >= (size_t)h->width + 4) ? 1 : 0;
    P.outKp = d_kp; P.outDesc = d_desc; P.outCount = d_counts; P.outCap = cap;
    int launches = 0;
    // Small host-path calls (one frame per call is what Frame::ExtractORB issues) are bound by launch overhead, not by
    // the kernels: their kernel sequence is captured once per batch size and replayed as one CUDA graph.  Every
    // parameter of such a call is the handle's own (staging slab in, output buffers out), so the replay is exact.
    const bool graphable = !h->profiling && staged && first == 0 && batch <= GRAPH_MAX_BATCH && d_images == h->dIn &&
                           d_kp == h->dOutKp && d_desc == h->dOutDesc && d_counts == h->dOutCount && cap == h->maxKp;
    if (graphable) {
        std::map<int, cudaGraphExec_t>::iterator it = h->graphs.find(batch);
        if (it == h->graphs.end()) {
            cudaGraph_t graph = nullptr;
            cudaGraphExec_t exec = nullptr;
            ORB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            int rc = launch_kernels(h, P, batch, st, launches);
            const cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (rc != ORBB200_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess) { set_error("graph capture failed: %s", cudaGetErrorString(ce)); return ORBB200_ECUDA; }
            ORB_CUDA(cudaGraphInstantiate(&exec, graph, 0));
            cudaGraphDestroy(graph);
            it = h->graphs.insert(std::make_pair(batch, exec)).first;
            h->graphLaunches[batch] = launches;
        } else {
            launches = h->graphLaunches[batch];
        }
        ORB_CUDA(cudaGraphLaunch(it->second, st));
    } else {
        int rc = launch_kernels(h, P, batch, st, launches);
        if (rc != ORBB200_OK) return rc;
    }
    h->lastLaunches = launches; h->lastBatch = batch;
    h->lastIn = d_images; h->lastInPitch = (int)stride; h->lastInFrameStride = (long long)frame_stride;
    return ORBB200_OK;
}

static int check_status(orbb200_extractor* h)
{
    int st = 0;
    ORB_CUDA(cudaMemcpyAsync(&st, h->P.status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    ORB_CUDA(cudaStreamSynchronize(h->stream));
    if (st) {
        cudaMemsetAsync(h->P.status, 0, sizeof(int), h->stream);
        set_error("device-side failure bits 0x%x (1=quadtree runaway, 2=candidate overflow, 4=keypoint overflow)", st);
        return ORBB200_ECUDA;
    }
    return ORBB200_OK;
}

extern "C" int orbb200_extract_device(orbb200_extractor* h, const uint8_t* d_images, int batch, size_t stride,
                                      size_t frame_stride, orbb200_keypoint* d_keypoints, uint8_t* d_descriptors,
                                      int32_t* d_counts, int cap)
{
    if (!h || !d_images) { set_error("null handle or image pointer"); return ORBB200_EINVAL; }
    if (batch < 1 || batch > h->maxBatch) { set_error("batch %d outside 1..%d", batch, h->maxBatch); return ORBB200_EINVAL; }
    if (stride < (size_t)h->width) { set_error("stride smaller than the frame width"); return ORBB200_EINVAL; }
    if (!d_keypoints && !d_descriptors && !d_counts) { d_keypoints = h->dOutKp; d_descriptors = h->dOutDesc; d_counts = h->dOutCount; cap = h->maxKp; }
    if (!d_keypoints || !d_descriptors || !d_counts) { set_error("output pointers must be all set or all NULL"); return ORBB200_EINVAL; }
    if (cap < h->maxKp) { set_error("cap %d < orbb200_extractor_max_keypoints() = %d", cap, h->maxKp); return ORBB200_ECAPACITY; }
    ORB_CUDA(cudaSetDevice(h->device));
    return enqueue(h, d_images, batch, stride, frame_stride, d_keypoints, d_descriptors, d_counts, cap);
}

extern "C" int orbb200_extractor_sync(orbb200_extractor* h)
{
    if (!h) { set_error("null handle"); return ORBB200_EINVAL; }
    ORB_CUDA(cudaSetDevice(h->device));
    return check_status(h);
}

extern "C" int orbb200_extractor_outputs(orbb200_extractor* h, orbb200_keypoint** d_keypoints, uint8_t** d_descriptors,
                                         int32_t** d_counts, int* cap)
{
    if (!h) { set_error("null handle"); return ORBB200_EINVAL; }
    if (d_keypoints) *d_keypoints = h->dOutKp;
    if (d_descriptors) *d_descriptors = h->dOutDesc;
    if (d_counts) *d_counts = h->dOutCount;
    if (cap) *cap = h->maxKp;
    return ORBB200_OK;
}

extern "C" int orbb200_extract_host_wait(orbb200_extractor* h)
{
    if (!h) { set_error("null handle"); return ORBB200_EINVAL; }
    if (!h->pending) return ORBB200_OK;
    ORB_CUDA(cudaSetDevice(h->device));
    h->pending = false;
    ORB_CUDA(cudaStreamSynchronize(h->copyOut));
    return check_status(h);
}

// `chunks` = pieces the batch is cut into so that copies and kernels of ONE call overlap (0 = choose).
static int extract_host_queue(orbb200_extractor* h, const uint8_t* images, int batch, size_t stride,
                              size_t frame_stride, orbb200_keypoint* keypoints, uint8_t* descriptors,
                              int32_t* counts, int cap, int chunks)
{
    if (!h || !images || !keypoints || !descriptors || !counts) { set_error("null argument"); return ORBB200_EINVAL; }
    if (batch < 1 || batch > h->maxBatch) { set_error("batch %d outside 1..%d", batch, h->maxBatch); return ORBB200_EINVAL; }
    if (stride < (size_t)h->width) { set_error("stride smaller than the frame width"); return ORBB200_EINVAL; }
    if (cap < h->maxKp) { set_error("cap %d < orbb200_extractor_max_keypoints() = %d", cap, h->maxKp); return ORBB200_ECAPACITY; }
    ORB_CUDA(cudaSetDevice(h->device));
    if (h->pending) {           // one call in flight per handle: its buffers are about to be overwritten
        int rc = orbb200_extract_host_wait(h);
        if (rc != ORBB200_OK) return rc;
    }
    // Chunked pipeline: while chunk c is in the kernels, chunk c+1 is on its way up and chunk c-1 on its way
    // down (three streams, events between them).  Small batches go through in one piece.
    int nchunk = chunks > 0 ? chunks : (batch >= 96 ? 3 : (batch >= 32 ? 2 : 1));   // measured on B200 + PCIe gen5: 2-3 chunks are best at batch 256
    if (chunks <= 0 && h->chunkOverride > 0) nchunk = std::min(batch, h->chunkOverride);
    const int cs = (batch + nchunk - 1) / nchunk;
    const int mk = h->maxKp;
    const size_t inFrameBytes = h->inPitch * (size_t)h->height;
    int launches = 0;
    for (int c = 0; c < nchunk; c++) {
        const int f0 = c * cs, n = std::min(cs, batch - f0);
        if (n <= 0) break;
        const uint8_t* src = images + (size_t)f0 * frame_stride;
        uint8_t* dIn = h->dIn + (size_t)f0 * inFrameBytes;
        if (frame_stride == stride * (size_t)h->height && stride == h->inPitch) {
            ORB_CUDA(cudaMemcpyAsync(dIn, src, inFrameBytes * n, cudaMemcpyHostToDevice, h->copyIn));       // one linear DMA
        } else if (frame_stride == stride * (size_t)h->height) {
            ORB_CUDA(cudaMemcpy2DAsync(dIn, h->inPitch, src, stride, h->width, (size_t)h->height * n, cudaMemcpyHostToDevice, h->copyIn));
        } else {
            for (int f = 0; f < n; f++)
                ORB_CUDA(cudaMemcpy2DAsync(dIn + (size_t)f * inFrameBytes, h->inPitch, src + f * frame_stride, stride, h->width,
                                           h->height, cudaMemcpyHostToDevice, h->copyIn));
        }
        ORB_CUDA(cudaEventRecord(h->evIn[c], h->copyIn));
        ORB_CUDA(cudaStreamWaitEvent(h->stream, h->evIn[c], 0));
        int rc = enqueue(h, dIn, n, h->inPitch, inFrameBytes, h->dOutKp + (size_t)f0 * mk, h->dOutDesc + (size_t)f0 * mk * 32,
                         h->dOutCount + f0, mk, f0);
        if (rc != ORBB200_OK) return rc;
        launches += h->lastLaunches;
        ORB_CUDA(cudaEventRecord(h->evDone[c], h->stream));
        ORB_CUDA(cudaStreamWaitEvent(h->copyOut, h->evDone[c], 0));
        ORB_CUDA(cudaMemcpyAsync(counts + f0, h->dOutCount + f0, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, h->copyOut));
        if (cap == mk) {
            ORB_CUDA(cudaMemcpyAsync(keypoints + (size_t)f0 * cap, h->dOutKp + (size_t)f0 * mk, (size_t)n * mk * sizeof(orbb200_keypoint),
                                     cudaMemcpyDeviceToHost, h->copyOut));
            ORB_CUDA(cudaMemcpyAsync(descriptors + (size_t)f0 * cap * 32, h->dOutDesc + (size_t)f0 * mk * 32, (size_t)n * mk * 32,
                                     cudaMemcpyDeviceToHost, h->copyOut));
        } else {
            ORB_CUDA(cudaMemcpy2DAsync(keypoints + (size_t)f0 * cap, (size_t)cap * sizeof(orbb200_keypoint), h->dOutKp + (size_t)f0 * mk,
                                       (size_t)mk * sizeof(orbb200_keypoint), (size_t)mk * sizeof(orbb200_keypoint), n,
                                       cudaMemcpyDeviceToHost, h->copyOut));
            ORB_CUDA(cudaMemcpy2DAsync(descriptors + (size_t)f0 * cap * 32, (size_t)cap * 32, h->dOutDesc + (size_t)f0 * mk * 32, (size_t)mk * 32,
                                       (size_t)mk * 32, n, cudaMemcpyDeviceToHost, h->copyOut));
        }
    }
    h->lastBatch = batch; h->lastIn = h->dIn; h->lastInPitch = (int)h->inPitch; h->lastInFrameStride = (long long)inFrameBytes;
    h->lastLaunches = launches;
    h->pending = true;
    return ORBB200_OK;
}

extern "C" int orbb200_extract_host(orbb200_extractor* h, const uint8_t* images, int batch, size_t stride,
                                    size_t frame_stride, orbb200_keypoint* keypoints, uint8_t* descriptors,
                                    int32_t* counts, int cap)
{
    int rc = extract_host_queue(h, images, batch, stride, frame_stride, keypoints, descriptors, counts, cap, 0);
    if (rc != ORBB200_OK) return rc;
    return orbb200_extract_host_wait(h);
}

// A streaming caller overlaps across calls (two or three handles in turn), where whole batches run best.
extern "C" int orbb200_extract_host_async(orbb200_extractor* h, const uint8_t* images, int batch, size_t stride,
                                          size_t frame_stride, orbb200_keypoint* keypoints, uint8_t* descriptors,
                                          int32_t* counts, int cap)
{
    return extract_host_queue(h, images, batch, stride, frame_stride, keypoints, descriptors, counts, cap, 1);
}

// ---- stage read-back -------------------------------------------------------------------
extern "C" int orbb200_extractor_get_level(orbb200_extractor* h, int frame, int level, int blurred, uint8_t* dst, size_t dst_stride)
{
    if (!h || !dst || level < 0 || level >= h->nlevels || frame < 0 || frame >= h->lastBatch) { set_error("bad argument"); return ORBB200_EINVAL; }
    ORB_CUDA(cudaSetDevice(h->device));
    const LevelGeo& g = h->P.lv[level];
    const uint8_t* src; size_t sp;
    if (blurred) { src = h->P.blur + (size_t)frame * h->P.blurFrameBytes + g.blurOff; sp = g.pitch; }
    else if (level == 0) { src = h->lastIn + (size_t)frame * h->lastInFrameStride; sp = h->lastInPitch; }
    else { src = h->P.pyr + (size_t)frame * h->P.pyrFrameBytes + g.pyrOff; sp = g.pitch; }
    ORB_CUDA(cudaStreamSynchronize(h->stream));
    ORB_CUDA(cudaMemcpy2D(dst, dst_stride, src, sp, g.w, g.h, cudaMemcpyDeviceToHost));
    return ORBB200_OK;
}

extern "C" int orbb200_extractor_pyramid_view(orbb200_extractor* h, orbb200_pyramid_view* v)
{
    if (!h || !v) { set_error("null argument"); return ORBB200_EINVAL; }
    if (h->lastBatch < 1 || !h->lastIn) { set_error("no extract call yet"); return ORBB200_EINVAL; }
    memset(v, 0, sizeof(*v));
    v->nlevels = h->nlevels;
    for (int l = 0; l < h->nlevels; l++) {
        const LevelGeo& g = h->P.lv[l];
        v->width[l] = g.w; v->height[l] = g.h;
        if (l == 0) { v->level[0] = h->lastIn; v->frame_stride[0] = (size_t)h->lastInFrameStride; v->pitch[0] = h->lastInPitch; }
        else { v->level[l] = h->P.pyr + g.pyrOff; v->frame_stride[l] = h->P.pyrFrameBytes; v->pitch[l] = g.pitch; }
    }
    return ORBB200_OK;
}

static int fetch_packed(orbb200_extractor* h, const uint32_t* dsrc, const int* dcount, int maxn, std::vector<uint32_t>& v)
{
    int n = 0;
    ORB_CUDA(cudaStreamSynchronize(h->stream));
    ORB_CUDA(cudaMemcpy(&n, dcount, sizeof(int), cudaMemcpyDeviceToHost));
    n = std::min(n, maxn);
    v.resize(n);
    if (n) ORB_CUDA(cudaMemcpy(v.data(), dsrc, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
    return ORBB200_OK;
}

extern "C" int orbb200_extractor_get_candidates(orbb200_extractor* h, int frame, int level, int32_t* x, int32_t* y,
                                                int32_t* score, int cap, int* n)
{
    if (!h || !n || level < 0 || level >= h->nlevels || frame < 0 || frame >= h->lastBatch) { set_error("bad argument"); return ORBB200_EINVAL; }
    ORB_CUDA(cudaSetDevice(h->device));
    const LevelGeo& g = h->P.lv[level];
    std::vector<uint32_t> v;
    int rc = fetch_packed(h, h->P.cand + (size_t)frame * h->P.candFrameCap + g.candOff, h->P.candCount + frame * h->nlevels + level, g.candCap, v);
    if (rc) return rc;
    // the device list is unordered; restore the reference's order: cells row-major, then row-major in the cell
    auto key = [&](uint32_t e) {
        const uint64_t px = e & 0xfff, py = (e >> 12) & 0xfff;
        return ((uint64_t)((py - 3) / g.hCell) << 40) | ((uint64_t)((px - 3) / g.wCell) << 28) | (py << 12) | px;
    };
    std::sort(v.begin(), v.end(), [&](uint32_t a, uint32_t b) { return key(a) < key(b); });
    *n = (int)v.size();
    if ((int)v.size() > cap) { set_error("candidate buffer too small (%zu)", v.size()); return ORBB200_ECAPACITY; }
    for (size_t i = 0; i < v.size(); i++) {
        if (x) x[i] = v[i] & 0xfff;
        if (y) y[i] = (v[i] >> 12) & 0xfff;
        if (score) score[i] = v[i] >> 24;
    }
    return ORBB200_OK;
}

extern "C" int orbb200_extractor_get_level_keypoints(orbb200_extractor* h, int frame, int level, int32_t* x, int32_t* y,
                                                     int32_t* score, int cap, int* n)
{
    if (!h || !n || level < 0 || level >= h->nlevels || frame < 0 || frame >= h->lastBatch) { set_error("bad argument"); return ORBB200_EINVAL; }
    ORB_CUDA(cudaSetDevice(h->device));
    const LevelGeo& g = h->P.lv[level];
    std::vector<uint32_t> v;
    int rc = fetch_packed(h, h->P.lkp + (size_t)frame * h->P.kpFrameCap + g.kpOff, h->P.lkpCount + frame * h->nlevels + level, g.kpCap, v);
    if (rc) return rc;
    *n = (int)v.size();
    if ((int)v.size() > cap) { set_error("keypoint buffer too small (%zu)", v.size()); return ORBB200_ECAPACITY; }
    for (size_t i = 0; i < v.size(); i++) {
        if (x) x[i] = (v[i] & 0xfff) + BORDER;
        if (y) y[i] = ((v[i] >> 12) & 0xfff) + BORDER;
        if (score) score[i] = v[i] >> 24;
    }
    return ORBB200_OK;
}
