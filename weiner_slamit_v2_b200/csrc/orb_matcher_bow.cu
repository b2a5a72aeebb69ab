// orb_matcher_bow.cu -- searches restricted to shared vocabulary nodes and the bag-of-words side of the front end:
//   SearchByBoW(pKF, F, vpMapPointMatches) S/ORBmatcher.cc:161-292, SearchByBoW(pKF1, pKF2, vpMatches12) :526-659,
//   SearchForTriangulation :661-827, MapPoint::ComputeDistinctiveDescriptors S/MapPoint.cc:248-313,
//   the DBoW2 transform of Frame::ComputeBoW S/Frame.cc:520-527 (Thirdparty/DBoW2 TemplatedVocabulary.h:1133-1266).
#include "matcher_common.cuh"

namespace orbb200 {

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
    ORB_CUDA(ensure_dynamic_smem((const void*)k_bow_assemble, m->device, sm));
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
