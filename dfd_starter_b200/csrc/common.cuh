// Shared helpers for the dfd_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>

#include "../../include/dfd_b200.h"

struct dfd_ctx {
    int device;
    int sm_count;
    int64_t launches;
    // sigma-scaled fp16 mirror of a noise table + fp16 scratch for theta (caller-owned memory registered by
    // dfd_table_build_scaled16; csrc/mlp_forward_direct.cu)
    const float* scaled_src;      // table->replicas the mirror was built from
    float scaled_sigma;
    void* scaled16;               // 8 replicas x scaled16_stride halves
    int64_t scaled16_stride;
    void* theta16;                // theta16_cap halves
    int64_t theta16_cap;
};

void dfd_set_error(const char* fmt, ...);

#define DFD_CHECK_ARG(cond, ...)          \
    do {                                  \
        if (!(cond)) {                    \
            dfd_set_error(__VA_ARGS__);   \
            return 1;                     \
        }                                 \
    } while (0)

#define DFD_CUDA(call)                                                                             \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            dfd_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return 2;                                                                              \
        }                                                                                          \
    } while (0)

// after a kernel launch: catch launch-configuration errors without synchronising
#define DFD_LAUNCHED(ctx)                                                                          \
    do {                                                                                           \
        (ctx)->launches++;                                                                         \
        cudaError_t e__ = cudaPeekAtLastError();                                                   \
        if (e__ != cudaSuccess) {                                                                  \
            dfd_set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return 3;                                                                              \
        }                                                                                          \
    } while (0)

static inline size_t dfd_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Programmatic dependent launch: the kernel may be scheduled while its predecessor on the stream is still draining
// (its CTAs start as the predecessor's CTAs exit), so launch latency and the dependent's prologue overlap the
// predecessor's tail.  Every kernel launched this way executes dfd_grid_dependency_wait() before it touches anything
// the predecessor wrote - the semantics stay those of an ordinary stream-ordered launch.  DFD_NO_PDL=1 disables it.
template <typename... KArgs, typename... Args>
static inline cudaError_t dfd_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                         Args... args) {
    static const bool off = getenv("DFD_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = off ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
__device__ __forceinline__ void dfd_grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// 16-byte streaming load: read-only path, do not allocate in L1 (rows are touched once per kernel)
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// base pointer of table[idx : idx+P] inside the 16-byte aligned replica (idx & 3)
__device__ __forceinline__ const float* table_row_ptr(const float* replicas, int64_t stride, int64_t idx) {
    const int64_t s = idx & 3;
    return replicas + s * stride + (idx - s);
}

// theta' = theta + sign*sigma*eps, bit-identical to numpy's `flat + sigma*eps`
// (worker/worker.py:28): product rounded to fp32, then the sum rounded; never an FMA.
__device__ __forceinline__ float perturb1(float theta, float sigma_signed, float eps) {
    return __fadd_rn(theta, __fmul_rn(sigma_signed, eps));
}
