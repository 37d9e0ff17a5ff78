// Shared helpers for the humaniflow_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cstdlib>
#include <cstdint>
#include <atomic>
#include "../../include/humaniflow_b200.h"

namespace hf {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define HF_CUDA(expr)                                                                       \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return hf::fail(HF_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,       \
                            cudaGetErrorString(_e));                                        \
    } while (0)

// first statement of a kernel launched with launch_pdl: wait until the predecessor grid has completed and its writes
// are visible (the successor is released implicitly when this grid's blocks exit)
#define HF_PDL_SYNC()                                              \
    do {                                                           \
        asm volatile("griddepcontrol.wait;" ::: "memory");         \
    } while (0)

#define HF_LAUNCH_CHECK()                                                                   \
    do {                                                                                    \
        hf::g_launches.fetch_add(1, std::memory_order_relaxed);                             \
        cudaError_t _e = cudaGetLastError();                                                \
        if (_e != cudaSuccess)                                                              \
            return hf::fail(HF_ERR_CUDA, "%s:%d launch -> %s", __FILE__, __LINE__,          \
                            cudaGetErrorString(_e));                                        \
    } while (0)

template <typename T>
inline int upload(T** dptr, const T* host, size_t n) {
    HF_CUDA(cudaMalloc((void**)dptr, n * sizeof(T)));
    HF_CUDA(cudaMemcpy(*dptr, host, n * sizeof(T), cudaMemcpyHostToDevice));
    return HF_OK;
}

inline int div_up(int a, int b) { return (a + b - 1) / b; }

// Launch with programmatic stream serialization: the kernel may be scheduled while its predecessor in the stream
// drains; every kernel launched this way starts with HF_PDL_SYNC() before it touches global memory.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("HF_NO_PDL") != nullptr;     // debugging aid
    cfg.attrs = at; cfg.numAttrs = no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace hf
