// int_peak.cu -- measures the integer-pipe peaks the matcher roofline is quoted against ("the builder
// must first measure the POPC peak with a microbenchmark"), and the register-only rate of three ways of computing a
// 256-bit Hamming distance, so the kernel design can be chosen from numbers:
//   popc      : independent POPC chains                     -> POPC lanes / clk / SM
//   lop3      : independent 3-input logic chains            -> LOP3 lanes / clk / SM
//   iadd3     : independent add chains
//   ham8      : 8 XOR + 8 POPC + adds                       (what hamming256 does)
//   ham6      : first-level carry-save adders, 6 POPC
//   ham4      : full carry-save tree (ones/twos/fours/eights), 4 POPC
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/int_peak tools/int_peak.cu ; prints one JSON line.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, ILP = 8;

__global__ void k_popc(uint32_t* out, uint32_t seed)
{
    uint32_t v[ILP], acc = 0;
    for (int i = 0; i < ILP; i++) v[i] = seed * (threadIdx.x + 1) + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) v[i] = __popc(v[i]) ^ seed;      // POPC + LOP (the LOP keeps the chain data dependent)
    }
    for (int i = 0; i < ILP; i++) acc += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void k_lop3(uint32_t* out, uint32_t seed)
{
    uint32_t v[ILP], acc = 0;
    for (int i = 0; i < ILP; i++) v[i] = seed * (threadIdx.x + 1) + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) v[i] = (v[i] & v[(i + 1) % ILP]) ^ seed;
    }
    for (int i = 0; i < ILP; i++) acc += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void k_iadd3(uint32_t* out, uint32_t seed)
{
    uint32_t v[ILP], acc = 0;
    for (int i = 0; i < ILP; i++) v[i] = seed * (threadIdx.x + 1) + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) v[i] = v[i] + v[(i + 1) % ILP] + seed;
    }
    for (int i = 0; i < ILP; i++) acc += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) { return a ^ b ^ c; }
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a | b)); }

template <int MODE> __device__ __forceinline__ int ham(const uint32_t* a, const uint32_t* b)
{
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = a[i] ^ b[i];
    if (MODE == 8) {
        int d = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) d += __popc(x[i]);
        return d;
    } else if (MODE == 6) {
        const uint32_t s1 = xor3(x[0], x[1], x[2]), c1 = maj3(x[0], x[1], x[2]);
        const uint32_t s2 = xor3(x[3], x[4], x[5]), c2 = maj3(x[3], x[4], x[5]);
        return __popc(s1) + __popc(s2) + __popc(x[6]) + __popc(x[7]) + 2 * (__popc(c1) + __popc(c2));
    } else {
        const uint32_t s1 = xor3(x[0], x[1], x[2]), c1 = maj3(x[0], x[1], x[2]);
        const uint32_t s2 = xor3(x[3], x[4], x[5]), c2 = maj3(x[3], x[4], x[5]);
        const uint32_t s3 = xor3(s1, s2, x[6]), c3 = maj3(s1, s2, x[6]);
        const uint32_t ones = s3 ^ x[7], c4 = s3 & x[7];
        const uint32_t t1 = xor3(c1, c2, c3), f1 = maj3(c1, c2, c3);
        const uint32_t twos = t1 ^ c4, f2 = t1 & c4;
        const uint32_t fours = f1 ^ f2, eights = f1 & f2;
        return __popc(ones) + 2 * __popc(twos) + 4 * __popc(fours) + 8 * __popc(eights);
    }
}

// one query per thread against candidates broadcast from shared memory, minimum kept: the skeleton of k_init_topk
template <int MODE> __global__ void k_ham(uint32_t* out, uint32_t seed)
{
    __shared__ uint32_t cand[64 * 8];
    for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) cand[i] = (i + 1) * 2654435761u ^ seed;
    __syncthreads();
    uint32_t a[8];
    for (int i = 0; i < 8; i++) a[i] = seed * (threadIdx.x + 1) + i * 40503u;
    int best = 1 << 30;
    for (int it = 0; it < ITERS / 64; it++) {
#pragma unroll 4
        for (int c = 0; c < 64; c++) {
            const int d = ham<MODE>(a, cand + c * 8);
            best = min(best, d + it);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = best;
}

template <typename F> static float time_ms(F launch)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 8, threads = 256;
    uint32_t* out; cudaMalloc(&out, (size_t)blocks * threads * 4);
    const double lanes = (double)blocks * threads;
    const double hz = clk_khz * 1e3;
    const float t_popc = time_ms([&] { k_popc<<<blocks, threads>>>(out, 12345u); });
    const float t_lop = time_ms([&] { k_lop3<<<blocks, threads>>>(out, 12345u); });
    const float t_add = time_ms([&] { k_iadd3<<<blocks, threads>>>(out, 12345u); });
    const float t_h8 = time_ms([&] { k_ham<8><<<blocks, threads>>>(out, 12345u); });
    const float t_h6 = time_ms([&] { k_ham<6><<<blocks, threads>>>(out, 12345u); });
    const float t_h4 = time_ms([&] { k_ham<4><<<blocks, threads>>>(out, 12345u); });
    const double ops = lanes * ITERS * ILP;
    const double evals = lanes * ITERS;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz\": %.0f, "
           "\"popc_per_clk_per_sm\": %.2f, \"popc_plus_lop_Gops\": %.1f, \"lop3_per_clk_per_sm\": %.2f, \"iadd3_per_clk_per_sm\": %.2f, "
           "\"hamming256_Geval_s\": {\"popc8\": %.1f, \"csa_popc6\": %.1f, \"csa_popc4\": %.1f}, "
           "\"how\": \"register-only chains, %d blocks x %d threads, best of 5, CUDA events; popc chain = POPC + one LOP per step\"}\n",
           p.name, sms, hz / 1e6, ops / (t_popc * 1e-3) / hz / sms, ops / (t_popc * 1e-3) / 1e9, ops / (t_lop * 1e-3) / hz / sms,
           ops / (t_add * 1e-3) / hz / sms, evals / (t_h8 * 1e-3) / 1e9, evals / (t_h6 * 1e-3) / 1e9, evals / (t_h4 * 1e-3) / 1e9,
           blocks, threads);
    return 0;
}
