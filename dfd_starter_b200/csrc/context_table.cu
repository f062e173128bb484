// Context management + the device noise table (utils/noise_sources.py:36-51).
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

void dfd_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int dfd_abi_version(void) { return DFD_ABI_VERSION; }
extern "C" const char* dfd_last_error(void) { return g_err; }

extern "C" int dfd_ctx_create(int device, dfd_ctx** out) {
    DFD_CHECK_ARG(out != nullptr, "dfd_ctx_create: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        dfd_set_error("dfd_ctx_create: no CUDA device (%s); this library has no CPU fallback",
                      e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return 2;
    }
    DFD_CHECK_ARG(device >= 0 && device < n, "dfd_ctx_create: device %d out of range (0..%d)", device, n - 1);
    cudaDeviceProp prop;
    DFD_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        dfd_set_error("dfd_ctx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                      prop.major, prop.minor);
        return 2;
    }
    DFD_CUDA(cudaSetDevice(device));
    dfd_ctx* c = new dfd_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->launches = 0;
    *out = c;
    return 0;
}

extern "C" int dfd_ctx_destroy(dfd_ctx* ctx) {
    delete ctx;
    return 0;
}
extern "C" int dfd_ctx_sm_count(const dfd_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
extern "C" int64_t dfd_ctx_launch_count(const dfd_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------
// table replicas + fp64 prefix sum of squares
// ---------------------------------------------------------------------------
extern "C" int64_t dfd_table_replica_stride(int64_t size) { return (size + 64 + 31) / 32 * 32; }

static const int SCAN_BLOCK = 4096;  // elements per scan block
static const int SCAN_THREADS = 256;

extern "C" size_t dfd_table_scratch_bytes(int64_t size) {
    int64_t nblk = (size + SCAN_BLOCK - 1) / SCAN_BLOCK;
    return (size_t)(nblk + 1) * sizeof(double);
}

__global__ void table_replicas_kernel(const float* __restrict__ table, int64_t size, float* __restrict__ replicas,
                                      int64_t stride) {
    // replica_s[j] = table[j + s]; zero padding past the end so vector loads that overhang a row stay defined
    const int64_t total = 4 * stride;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = i / stride, j = i - s * stride;
        const int64_t src = j + s;
        replicas[i] = src < size ? table[src] : 0.0f;
    }
}

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    if (w == 0) {
        t = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) sh[0] = t;
    }
    __syncthreads();
    t = sh[0];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(SCAN_THREADS) sq_block_sums_kernel(const float* __restrict__ table, int64_t size,
                                                                     double* __restrict__ block_sums) {
    __shared__ double sh[32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK;
    double acc = 0.0;
    for (int i = threadIdx.x; i < SCAN_BLOCK; i += SCAN_THREADS) {
        const int64_t j = base + i;
        if (j < size) {
            const double v = (double)table[j];
            acc += v * v;
        }
    }
    const double t = block_sum_d(acc, sh);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = t;
}

// single block: exclusive scan of block sums, in place (nblk is a few thousand)
__global__ void __launch_bounds__(1024) scan_block_sums_kernel(double* __restrict__ block_sums, int64_t nblk) {
    __shared__ double sh[32];
    __shared__ double carry_s;
    if (threadIdx.x == 0) carry_s = 0.0;
    __syncthreads();
    for (int64_t base = 0; base < nblk; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const double v = i < nblk ? block_sums[i] : 0.0;
        // inclusive scan within the warp
        double x = v;
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) sh[w] = x;
        __syncthreads();
        if (w == 0) {
            double s = sh[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            sh[lane] = s;
        }
        __syncthreads();
        const double warp_off = w > 0 ? sh[w - 1] : 0.0;
        const double carry = carry_s;
        if (i < nblk) block_sums[i] = carry + warp_off + x - v;  // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_off + x;
        __syncthreads();
    }
}

// prefix[i] = sum_{j<i} table[j]^2 ; one CTA per scan block, sequential 16-element runs per thread
__global__ void __launch_bounds__(SCAN_THREADS) sq_prefix_kernel(const float* __restrict__ table, int64_t size,
                                                                 const double* __restrict__ block_off,
                                                                 double* __restrict__ prefix) {
    constexpr int PER = SCAN_BLOCK / SCAN_THREADS;  // 16
    __shared__ double sh[32];
    __shared__ double warp_tot[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * PER;
    double v[PER];
    double run = 0.0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int64_t j = base + k;
        const double x = j < size ? (double)table[j] : 0.0;
        run += x * x;
        v[k] = run;  // inclusive within the thread
    }
    // exclusive scan of per-thread totals across the block
    double x = run;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0.0;
        for (int i = 0; i < SCAN_THREADS / 32; ++i) {
            sh[i] = c;
            c += warp_tot[i];
        }
    }
    __syncthreads();
    const double off = block_off[blockIdx.x] + sh[w] + (x - run);
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int64_t j = base + k;
        if (j < size) prefix[j + 1] = off + v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) prefix[0] = 0.0;
}

extern "C" int dfd_table_build(dfd_ctx* ctx, const float* table_dev, int64_t size, float* replicas,
                               int64_t replica_stride, double* prefix_sq, void* scratch, size_t scratch_bytes,
                               dfd_stream stream) {
    DFD_CHECK_ARG(ctx && table_dev && replicas && prefix_sq && scratch, "dfd_table_build: NULL argument");
    DFD_CHECK_ARG(size > 0, "dfd_table_build: size must be positive");
    DFD_CHECK_ARG(replica_stride >= size + 64 && replica_stride % 32 == 0,
                  "dfd_table_build: replica_stride %lld must be a multiple of 32 and >= size+64", (long long)replica_stride);
    DFD_CHECK_ARG(((uintptr_t)replicas & 15) == 0, "dfd_table_build: replicas must be 16-byte aligned");
    DFD_CHECK_ARG(scratch_bytes >= dfd_table_scratch_bytes(size), "dfd_table_build: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nblk = (size + SCAN_BLOCK - 1) / SCAN_BLOCK;
    double* block_sums = (double*)scratch;
    table_replicas_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(table_dev, size, replicas, replica_stride);
    DFD_LAUNCHED(ctx);
    sq_block_sums_kernel<<<(unsigned)nblk, SCAN_THREADS, 0, st>>>(table_dev, size, block_sums);
    DFD_LAUNCHED(ctx);
    scan_block_sums_kernel<<<1, 1024, 0, st>>>(block_sums, nblk);
    DFD_LAUNCHED(ctx);
    sq_prefix_kernel<<<(unsigned)nblk, SCAN_THREADS, 0, st>>>(table_dev, size, block_sums, prefix_sq);
    DFD_LAUNCHED(ctx);
    return 0;
}

// ---------------------------------------------------------------------------
// perturbation (worker/worker.py:28), materialised
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) perturb_members_kernel(const float* __restrict__ replicas, int64_t stride,
                                                              const float* __restrict__ theta, int64_t P,
                                                              const int64_t* __restrict__ idx,
                                                              const int8_t* __restrict__ sign, float sigma,
                                                              float* __restrict__ out, int64_t out_stride) {
    const int m = blockIdx.y;
    const float sg = sigma * (float)sign[m];
    const float* row = table_row_ptr(replicas, stride, idx[m]);
    float* o = out + (int64_t)m * out_stride;
    const int64_t nvec = P >> 2;
    const bool vec_ok = ((((uintptr_t)theta) | ((uintptr_t)o)) & 15) == 0;
    if (vec_ok) {
        for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
            const float4 e = ldg_stream_f4(row + 4 * v);
            const float4 t = *reinterpret_cast<const float4*>(theta + 4 * v);
            float4 r;
            r.x = perturb1(t.x, sg, e.x);
            r.y = perturb1(t.y, sg, e.y);
            r.z = perturb1(t.z, sg, e.z);
            r.w = perturb1(t.w, sg, e.w);
            *reinterpret_cast<float4*>(o + 4 * v) = r;
        }
        for (int64_t p = (nvec << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P;
             p += (int64_t)gridDim.x * blockDim.x)
            o[p] = perturb1(theta[p], sg, row[p]);
    } else {
        for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x)
            o[p] = perturb1(theta[p], sg, row[p]);
    }
}

extern "C" int dfd_perturb_members(dfd_ctx* ctx, const dfd_table* table, const float* theta, int64_t n_params,
                                   const int64_t* idx, const int8_t* sign, int n_members, float sigma, float* out,
                                   int64_t out_stride, dfd_stream stream) {
    DFD_CHECK_ARG(ctx && table && theta && idx && sign && out, "dfd_perturb_members: NULL argument");
    DFD_CHECK_ARG(n_params > 0 && n_params < table->size, "dfd_perturb_members: n_params %lld vs table size %lld",
                  (long long)n_params, (long long)table->size);
    DFD_CHECK_ARG(out_stride >= n_params, "dfd_perturb_members: out_stride < n_params");
    if (n_members == 0) return 0;
    DFD_CHECK_ARG(n_members > 0 && n_members <= 65535, "dfd_perturb_members: n_members out of range");
    int gx = (int)((n_params / 4 + 255) / 256);
    if (gx < 1) gx = 1;
    if (gx > 64) gx = 64;
    dim3 grid(gx, n_members);
    perturb_members_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(table->replicas, table->replica_stride, theta,
                                                                    n_params, idx, sign, sigma, out, out_stride);
    DFD_LAUNCHED(ctx);
    return 0;
}

// ---------------------------------------------------------------------------
// small host -> device staging that does not use the copy engine
// ---------------------------------------------------------------------------
// The learner's per-step batch (rewards | indices | history rows | signs, ~21 bytes per return) is tens of KB.  As a
// cudaMemcpyAsync it queues on the host-to-device copy engine BEHIND whatever large upload is in flight there (the
// next step's observations, 18 MB = 0.33 ms over PCIe 5 x16), which put that whole transfer on the learner step's
// critical path.  Here the SMs read the pinned buffer directly through its device alias (unified addressing) and
// write the device copy: a few PCIe read round trips, independent of the copy engine's queue.  The same holds for the
// small results going back (rewards, update size, theta of a short parameter vector): measured on B200, a 16 KB
// device-to-host cudaMemcpyAsync issued while an 18 MB host-to-device copy was in flight completed only after it; here
// the SMs write the pinned destination directly (posted PCIe writes).
__global__ void __launch_bounds__(256) host_stage_kernel(const int4* __restrict__ src, int4* __restrict__ dst, int64_t n16,
                                                         const unsigned char* __restrict__ src_tail,
                                                         unsigned char* __restrict__ dst_tail, int n_tail) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
        int4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(src + i));
        dst[i] = v;
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < n_tail) dst_tail[threadIdx.x] = src_tail[threadIdx.x];
}

// device-visible address of a device pointer or of PINNED host memory (its alias under unified addressing)
static int dfd_device_view(const void* p, void** out) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        dfd_set_error("dfd_host_stage: cudaPointerGetAttributes: %s", cudaGetErrorString(e));
        return 2;
    }
    if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) {
        *out = const_cast<void*>(p);
        return 0;
    }
    if (a.type == cudaMemoryTypeHost && a.devicePointer) {
        *out = a.devicePointer;
        return 0;
    }
    dfd_set_error("dfd_host_stage: %p is neither device memory nor pinned (page-locked) host memory", p);
    return 1;
}

extern "C" int dfd_host_stage(dfd_ctx* ctx, const void* src, void* dst, size_t bytes, dfd_stream stream) {
    DFD_CHECK_ARG(ctx && src && dst, "dfd_host_stage: NULL argument");
    DFD_CHECK_ARG((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "dfd_host_stage: buffers must be 16-byte aligned");
    if (bytes == 0) return 0;
    void *s = nullptr, *d = nullptr;
    int rc = dfd_device_view(src, &s);
    if (rc) return rc;
    rc = dfd_device_view(dst, &d);
    if (rc) return rc;
    const int64_t n16 = (int64_t)(bytes / 16);
    const int n_tail = (int)(bytes % 16);
    int grid = (int)((n16 + 255) / 256);
    grid = grid < 1 ? 1 : (grid > 2 * ctx->sm_count ? 2 * ctx->sm_count : grid);
    host_stage_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const int4*)s, (int4*)d, n16,
                                                              (const unsigned char*)s + n16 * 16,
                                                              (unsigned char*)d + n16 * 16, n_tail);
    DFD_LAUNCHED(ctx);
    return 0;
}
