// Shared helpers for libfrustum_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <utility>

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

// Programmatic dependent launch (PDL): every kernel of the forward chain is launched with the
// programmatic-stream-serialization attribute, executes `griddepcontrol.wait` before its first
// global-memory access (full completion + memory flush of the previous grid) and releases its own
// dependents right after, so launch latency / CTA rasterisation / barrier+TMEM prologues of kernel
// N+1 overlap the tail of kernel N.  Works unchanged under CUDA-graph capture.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
// Dependents are released right after the wait (shortest chain).  Releasing them only when a CTA has issued its
// last tile was measured within +-1 % on the car bench and dropped.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }

// A/B knob: launch priority of a kernel class (kept by the graph node when captured).  `FCN_PRIO_PN` /
// `FCN_PRIO_CONV` = integer in the device's stream-priority range (lower = scheduled first); unset = no attribute.
// With several forwards in flight the block scheduler then places e.g. the persistent PointNet CTAs before the
// conv CTAs of the other forwards instead of in arrival order.
constexpr int FCN_NO_PRIORITY = -1000000;
inline int env_priority(const char *name) {
    const char *v = getenv(name);
    return (v != nullptr && *v != 0) ? atoi(v) : FCN_NO_PRIORITY;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_prio(int priority, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                   cudaStream_t stream, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    static const bool no_pdl = getenv("FCN_NO_PDL") != nullptr;   // diagnostics
    unsigned n = 0;
    if (!no_pdl) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (priority != FCN_NO_PRIORITY) {
        attr[n].id = cudaLaunchAttributePriority;
        attr[n].val.priority = priority;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args &&...args) {
    return launch_pdl_prio(FCN_NO_PRIORITY, kern, grid, block, smem, stream, std::forward<Args>(args)...);
}

}  // namespace fcn
