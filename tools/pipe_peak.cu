// pipe_peak.cu -- register-only issue-rate microbenchmarks for the instruction classes the extractor kernels are made of,
// alone and in pairs, so that "which pipe does this run on, and do the two overlap" is answered by a number on B200
// instead of by folklore.  Each kernel runs ILP independent dependency chains of one instruction (or an A,B mix) per
// thread; the result is lanes / clock / SM.  Two ops that sit on different pipes show a mixed rate above either alone.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_peak tools/pipe_peak.cu ; prints one JSON object.
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int ITERS = 2048, ILP = 8;

#define CHAIN_KERNEL(name, BODY)                                                        \
    __global__ void name(uint32_t* out, uint32_t seed)                                  \
    {                                                                                   \
        uint32_t v[ILP], acc = 0;                                                       \
        const uint32_t s = seed ^ 0x64036405u, s2 = seed * 3u + 0x64116407u;            \
        for (int i = 0; i < ILP; i++) v[i] = (seed * (threadIdx.x + 1) + i) & 0x03ff03ffu | 0x64006400u; \
        for (int it = 0; it < ITERS; it++) {                                            \
            _Pragma("unroll") for (int i = 0; i < ILP; i++) { BODY }                    \
        }                                                                               \
        for (int i = 0; i < ILP; i++) acc += v[i];                                      \
        out[blockIdx.x * blockDim.x + threadIdx.x] = acc + s2;                          \
    }

__device__ __forceinline__ uint32_t hmin2_u(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hmax2_u(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hadd2_u(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hfma2_u(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t hset2_gt(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("set.gt.u32.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));      // 0xffff per half where a > b
    return d;
}
__device__ __forceinline__ float fmin3(float a, float b, float c)
{
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

CHAIN_KERNEL(k_vmin2, v[i] = __vminu2(v[i], v[(i + 1) % ILP] ^ s);)            // VIMNMX.U16x2 (+LOP3)
CHAIN_KERNEL(k_vmin2_pure, v[i] = __vminu2(__vmaxu2(v[i], s), s2);)             // two VIMNMX.U16x2
CHAIN_KERNEL(k_vmin3, v[i] = __vimin3_u16x2(v[i], v[(i + 1) % ILP], s);)        // VIMNMX3.U16x2
CHAIN_KERNEL(k_vmin3max3, v[i] = __vimax3_u16x2(__vimin3_u16x2(v[i], v[(i + 1) % ILP], s), s2, v[(i + 2) % ILP]);)
CHAIN_KERNEL(k_hmin2, v[i] = hmin2_u(hmax2_u(v[i], s), s2);)                    // two HMNMX2
CHAIN_KERNEL(k_hadd2, v[i] = hadd2_u(v[i], s);)
CHAIN_KERNEL(k_hfma2, v[i] = hfma2_u(v[i], s, s2);)
CHAIN_KERNEL(k_hset2, v[i] = hset2_gt(v[i], s) | 0x64006400u;)                  // HSET2 + LOP3
CHAIN_KERNEL(k_fmnmx, v[i] = __float_as_uint(fminf(fmaxf(__uint_as_float(v[i]), __uint_as_float(s)), __uint_as_float(s2)));)
CHAIN_KERNEL(k_fmnmx3, v[i] = __float_as_uint(fmin3(__uint_as_float(v[i]), __uint_as_float(v[(i + 1) % ILP]), __uint_as_float(s)));)
CHAIN_KERNEL(k_ffma, v[i] = __float_as_uint(__fmaf_rn(__uint_as_float(v[i]), 1.0000001f, __uint_as_float(s)));)
CHAIN_KERNEL(k_prmt, v[i] = __byte_perm(v[i], v[(i + 1) % ILP], 0x5432);)
CHAIN_KERNEL(k_lop3, v[i] = (v[i] & v[(i + 1) % ILP]) ^ s;)
CHAIN_KERNEL(k_iadd3, v[i] = v[i] + v[(i + 1) % ILP] + s;)
CHAIN_KERNEL(k_imad, v[i] = v[i] * s + v[(i + 1) % ILP];)
CHAIN_KERNEL(k_shf, v[i] = __funnelshift_r(v[i], v[(i + 1) % ILP], 8);)
CHAIN_KERNEL(k_dp4a, v[i] = __dp4a(v[i], s, v[(i + 1) % ILP]);)
CHAIN_KERNEL(k_dp2a, v[i] = __dp2a_lo(v[i], s, v[(i + 1) % ILP]);)
CHAIN_KERNEL(k_popc, v[i] = __popc(v[i]) ^ s;)
CHAIN_KERNEL(k_vabsdiff, v[i] = __vabsdiffu2(v[i], s);)
CHAIN_KERNEL(k_viaddmax, v[i] = __viaddmax_u16x2(v[i], s, s2);)
CHAIN_KERNEL(k_vsetgt2, v[i] = __vcmpgtu2(v[i], s) | 0x64006400u;)
CHAIN_KERNEL(k_ballot, v[i] = __ballot_sync(0xffffffffu, v[i] & 1) + s;)
CHAIN_KERNEL(k_shfl, v[i] = __shfl_xor_sync(0xffffffffu, v[i], 1) + s;)
CHAIN_KERNEL(k_isetp_sel, v[i] = (v[i] > s) ? v[(i + 1) % ILP] : s2;)
// mixes: one op of each kind per chain step
CHAIN_KERNEL(k_mix_vmin3_hmin2, v[i] = (i & 1) ? __vimin3_u16x2(v[i], v[(i + 1) % ILP], s) : hmin2_u(hmax2_u(v[i], s), s2);)   // 1 VIMNMX3 : 2 HMNMX2
CHAIN_KERNEL(k_mix_vmin3_imad, v[i] = (i & 1) ? __vimin3_u16x2(v[i], v[(i + 1) % ILP], s) : v[i] * s + v[(i + 1) % ILP];)
CHAIN_KERNEL(k_mix_vmin3_ffma, v[i] = (i & 1) ? __vimin3_u16x2(v[i], v[(i + 1) % ILP], s) : __float_as_uint(__fmaf_rn(__uint_as_float(v[i]), 1.0000001f, __uint_as_float(s)));)
CHAIN_KERNEL(k_mix_prmt_hmin2, v[i] = (i & 1) ? __byte_perm(v[i], v[(i + 1) % ILP], 0x5432) : hmin2_u(hmax2_u(v[i], s), s2);)
CHAIN_KERNEL(k_mix_prmt_imad, v[i] = (i & 1) ? __byte_perm(v[i], v[(i + 1) % ILP], 0x5432) : v[i] * s + v[(i + 1) % ILP];)
CHAIN_KERNEL(k_mix_vmin3_hset2, v[i] = (i & 1) ? __vimin3_u16x2(v[i], v[(i + 1) % ILP], s) : (hset2_gt(v[i], s) | 0x64006400u);)
CHAIN_KERNEL(k_mix_dp4a_imad, v[i] = (i & 1) ? __dp4a(v[i], s, v[(i + 1) % ILP]) : v[i] * s + v[(i + 1) % ILP];)
CHAIN_KERNEL(k_mix_vmin3_fmnmx, v[i] = (i & 1) ? __vimin3_u16x2(v[i], v[(i + 1) % ILP], s) : __float_as_uint(fminf(fmaxf(__uint_as_float(v[i]), __uint_as_float(s)), __uint_as_float(s2)));)

// shared-memory load rates (conflict-free): bytes / clock / SM and instructions
template <int W> __global__ void k_lds(uint32_t* out, uint32_t seed)
{
    __shared__ __align__(16) uint32_t sm[256 * 4 + 64];
    for (int i = threadIdx.x; i < 256 * 4 + 64; i += blockDim.x) sm[i] = i * seed;
    __syncthreads();
    uint32_t acc = 0;
    int off = (threadIdx.x & 31) * W;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            const int o = (off + i * 32 * W) & (256 * 4 - 1);
            if (W == 1) acc += sm[o];
            else if (W == 2) { const uint2 t = *reinterpret_cast<const uint2*>(sm + o); acc += t.x ^ t.y; }
            else { const uint4 t = *reinterpret_cast<const uint4*>(sm + o); acc += t.x ^ t.y ^ t.z ^ t.w; }
        }
        off = (off + (acc & 4) * W) & (256 * 4 - 1);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
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
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return best;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 8, threads = 256;
    uint32_t* out; cudaMalloc(&out, (size_t)blocks * threads * 4);
    const double lanes = (double)blocks * threads, hz = clk_khz * 1e3;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz\": %.0f, \"unit\": \"chain steps (lanes) per clock per SM; ops_per_step says how many instructions one step is\"", p.name, sms, hz / 1e6);
#define RUN(k, ops)                                                                              \
    {                                                                                            \
        const float t = time_ms([&] { k<<<blocks, threads>>>(out, 12345u); });                   \
        printf(", \"%s\": {\"steps_per_clk_sm\": %.2f, \"ops_per_step\": \"%s\"}", #k + 2, lanes * ITERS * ILP / (t * 1e-3) / hz / sms, ops); \
    }
    RUN(k_vmin2, "VIMNMX.U16x2 + LOP3")
    RUN(k_vmin2_pure, "2 VIMNMX.U16x2")
    RUN(k_vmin3, "VIMNMX3.U16x2")
    RUN(k_vmin3max3, "2 VIMNMX3.U16x2")
    RUN(k_hmin2, "2 HMNMX2")
    RUN(k_hadd2, "HADD2")
    RUN(k_hfma2, "HFMA2")
    RUN(k_hset2, "HSET2 + LOP3")
    RUN(k_fmnmx, "2 FMNMX")
    RUN(k_fmnmx3, "FMNMX3")
    RUN(k_ffma, "FFMA")
    RUN(k_prmt, "PRMT")
    RUN(k_lop3, "LOP3")
    RUN(k_iadd3, "IADD3")
    RUN(k_imad, "IMAD")
    RUN(k_shf, "SHF")
    RUN(k_dp4a, "IDP.4A")
    RUN(k_dp2a, "IDP.2A")
    RUN(k_popc, "POPC + LOP3")
    RUN(k_vabsdiff, "vabsdiffu2")
    RUN(k_viaddmax, "VIADDMNMX.U16x2")
    RUN(k_vsetgt2, "vcmpgtu2 + LOP3")
    RUN(k_ballot, "VOTE + IADD")
    RUN(k_shfl, "SHFL + IADD")
    RUN(k_isetp_sel, "ISETP + SEL")
    RUN(k_mix_vmin3_hmin2, "half the chains VIMNMX3, half 2 HMNMX2")
    RUN(k_mix_vmin3_imad, "half VIMNMX3, half IMAD")
    RUN(k_mix_vmin3_ffma, "half VIMNMX3, half FFMA")
    RUN(k_mix_prmt_hmin2, "half PRMT, half 2 HMNMX2")
    RUN(k_mix_prmt_imad, "half PRMT, half IMAD")
    RUN(k_mix_vmin3_hset2, "half VIMNMX3, half HSET2 + LOP3")
    RUN(k_mix_dp4a_imad, "half IDP.4A, half IMAD")
    RUN(k_mix_vmin3_fmnmx, "half VIMNMX3, half 2 FMNMX")
    RUN(k_lds<1>, "LDS.32")
    RUN(k_lds<2>, "LDS.64")
    RUN(k_lds<4>, "LDS.128")
    printf("}\n");
    return 0;
}
