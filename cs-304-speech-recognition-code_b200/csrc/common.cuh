// Shared helpers for the loe_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math_constants.h>
#include "../../include/loe_b200.h"

namespace loe {

void set_error(const char* fmt, ...);

inline int check_cuda(cudaError_t e, const char* what) {
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return LOE_ERR_CUDA;
    }
    return LOE_OK;
}

#define LOE_CUDA(call)                                      \
    do {                                                    \
        int _s = ::loe::check_cuda((call), #call);          \
        if (_s != LOE_OK) return _s;                        \
    } while (0)

#define LOE_LAUNCH_CHECK(name) LOE_CUDA(cudaGetLastError())

__device__ __forceinline__ float neg_inf() { return -CUDART_INF_F; }

}  // namespace loe
