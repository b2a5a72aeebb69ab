// common.cuh -- error plumbing shared by the extractor and matcher translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
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

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace orbb200
