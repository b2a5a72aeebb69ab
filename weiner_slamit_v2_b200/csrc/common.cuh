// common.cuh -- error plumbing shared by the extractor and matcher translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <map>
#include <mutex>
#include <utility>
#include "../../include/orb_b200.h"

namespace orbb200 {

void set_error(const char* fmt, ...);

// Evaluate a CUDA runtime call; on failure record the message and return ORBB200_ECUDA.
#define ORB_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (call);                                                                \
        if (_e != cudaSuccess) {                                                                \
            orbb200::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,             \
                               cudaGetErrorString(_e));                                         \
            return ORBB200_ECUDA;                                                               \
        }                                                                                       \
    } while (0)

#define ORB_CHECK_LAUNCH(name)                                                                  \
    do {                                                                                        \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess) {                                                                \
            orbb200::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));        \
            return ORBB200_ECUDA;                                                               \
        }                                                                                       \
    } while (0)

// Raises a kernel's dynamic shared-memory limit to at least `bytes` on `device` (the current device).  The limit is state
// of the (function, device) pair, shared by every handle and every host thread of the process, so it only ever grows and
// it changes under a lock: a handle created later with a smaller need can never lower it under a launch of an older one.
inline cudaError_t ensure_dynamic_smem(const void* kernel, int device, size_t bytes)
{
    static std::mutex mtx;
    static std::map<std::pair<const void*, int>, size_t> limit;
    if (bytes <= 48 * 1024) return cudaSuccess;
    std::lock_guard<std::mutex> lock(mtx);
    size_t& cur = limit[std::make_pair(kernel, device)];
    if (bytes <= cur) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace orbb200
