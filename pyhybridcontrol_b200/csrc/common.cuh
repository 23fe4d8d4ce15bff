// common.cuh -- shared helpers for libhmpc (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/hmpc.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libhmpc is written for sm_100a (B200) only"
#endif

namespace hmpc {

extern thread_local char g_last_error[256];

inline int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
    return HMPC_ERR_CUDA;
}

#define HMPC_CUDA_TRY(expr)                                              \
    do {                                                                 \
        cudaError_t e__ = (expr);                                        \
        if (e__ != cudaSuccess) return ::hmpc::cuda_fail(e__, #expr);    \
    } while (0)

#define HMPC_LAUNCH_CHECK(name)                                          \
    do {                                                                 \
        cudaError_t e__ = cudaGetLastError();                            \
        if (e__ != cudaSuccess) return ::hmpc::cuda_fail(e__, name);     \
    } while (0)

constexpr int kNumSM = 148;  // B200: 2 dies x 74 SMs

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace hmpc
