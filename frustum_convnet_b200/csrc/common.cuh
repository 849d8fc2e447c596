// Shared helpers for libfrustum_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/frustum_b200.h"

namespace fcn {

extern thread_local char g_err[512];

inline int invalid(const char *fn, const char *msg) {
    snprintf(g_err, sizeof(g_err), "%s: invalid argument: %s", fn, msg);
    return FCN_ERR_INVALID;
}
inline int cuda_fail(const char *fn, cudaError_t e) {
    snprintf(g_err, sizeof(g_err), "%s: CUDA error: %s", fn, cudaGetErrorString(e));
    return FCN_ERR_CUDA;
}

#define FCN_REQUIRE(cond, msg)                                  \
    do {                                                        \
        if (!(cond)) return fcn::invalid(__func__, msg);        \
    } while (0)

#define FCN_CUDA(call)                                              \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return fcn::cuda_fail(__func__, e__); \
    } while (0)

#define FCN_LAUNCH_CHECK()                                          \
    do {                                                            \
        cudaError_t e__ = cudaGetLastError();                       \
        if (e__ != cudaSuccess) return fcn::cuda_fail(__func__, e__); \
    } while (0)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

int sm_count();

}  // namespace fcn
