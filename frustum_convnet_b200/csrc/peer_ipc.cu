// Multi-GPU plumbing of the inference result exchange: CUDA-IPC export / import of a
// caller-owned device buffer, so that the heads epilogue of the persistent FCN kernel of one process can store
// straight into the gather buffer of another process's GPU over NVLink (fcn_mega_args.outs[1..]).
// The library still allocates nothing: the buffer is the caller's (a torch tensor), the import only maps it.
#include <cuda.h>

#include "common.cuh"

using namespace fcn;

typedef CUresult (*GetAddressRangeFn)(CUdeviceptr *, size_t *, CUdeviceptr);

extern "C" int fcn_ipc_export(const void *dev_ptr, void *handle_64B, long long *offset_bytes) {
    FCN_REQUIRE(dev_ptr && handle_64B && offset_bytes, "NULL pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    static GetAddressRangeFn range_fn = nullptr;
    if (range_fn == nullptr) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return fcn::invalid(__func__, "cuMemGetAddressRange is not available in this driver");
        range_fn = (GetAddressRangeFn)ptr;
    }
    CUdeviceptr base = 0;
    size_t size = 0;
    if (range_fn(&base, &size, (CUdeviceptr)dev_ptr) != CUDA_SUCCESS)
        return fcn::invalid(__func__, "not a device allocation of this context");
    cudaIpcMemHandle_t h;
    FCN_CUDA(cudaIpcGetMemHandle(&h, (void *)base));      // handle of the whole allocation the pointer lives in
    memcpy(handle_64B, &h, 64);
    *offset_bytes = (long long)((CUdeviceptr)dev_ptr - base);
    return FCN_OK;
}

// Maps the exporter's allocation into THIS process for the CURRENT device (peer access is enabled lazily by the
// driver: cudaIpcMemLazyEnablePeerAccess).  One open per handle and process; close with fcn_ipc_close.
extern "C" int fcn_ipc_open(const void *handle_64B, void **base_out) {
    FCN_REQUIRE(handle_64B && base_out, "NULL pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_64B, 64);
    FCN_CUDA(cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess));
    return FCN_OK;
}

extern "C" int fcn_ipc_close(void *base) {
    if (base == nullptr) return FCN_OK;
    FCN_CUDA(cudaIpcCloseMemHandle(base));
    return FCN_OK;
}
