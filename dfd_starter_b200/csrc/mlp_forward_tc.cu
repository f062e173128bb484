// Per-member perturbed MLP forward on the 5th-gen tensor cores (tcgen05 + TMEM), tf32 operands,
// fp32 accumulate.  MuJoCo head only (policies/mujoco.py:35-41 + utils/torch_helpers.py:20-25).
//
// One CTA = one (member, 128-observation tile).  out = act(X . W^T + b) per layer with
//   A operand = activations  [M = 128 observations][K]   (K-major, shared memory)
//   B operand = weights      [N = out features][K]       (K-major, shared memory; W is (out,in) row-major
//                                                         in the flat vector, so K is already contiguous)
//   D         = accumulator  [128 lanes][N columns]      (TMEM)
// The B operand cannot come from TMA: it is theta_tile + sign*sigma*eps_tile (worker/worker.py:28).
// All threads build it in shared memory in the UMMA canonical no-swizzle K-major layout
//   (8-row x 16-byte core matrices; LBO = 128 B between K-adjacent cores, SBO = KC*32 B between row groups)
// with the SAME two-rounding perturbation as the exact path, then round the operand to tf32 (rna).
// One elected thread issues tcgen05.mma (kind::tf32, M=128, N, K=8 per instruction) and commits to an
// mbarrier; the epilogue reads TMEM with tcgen05.ld, adds the perturbed bias, applies tanh and writes the
// next layer's A operand straight into the canonical layout (or the head to global memory).
#include "common.cuh"

namespace {

constexpr int TC_THREADS = 256;
constexpr int TC_KC = 64;  // K columns staged per chunk

struct TcLayout {
    int K0, K0p, N1, N2, nout, N3, A;
    int w_off[3], b_off[3], kin[3], kpad[3], nreal[3], npad[3];
    int64_t P;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// no-swizzle K-major canonical layout used below, offsets in floats for element (r, k) of a [rows x kc] tile:
//   (r >> 3) * (kc * 8) + (k >> 2) * 32 + (r & 7) * 4 + (k & 3)

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version for sm_100
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

// kind::tf32, fp32 accumulate, A and B K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();   // a lost commit must fault, not hang the GPU
    } while (!ok);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// tanh to ~1e-6 absolute: 1 - 2/(exp(2x)+1) with ex2.approx / rcp.approx (the 1-instruction tanh.approx is 5e-4)
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = __expf(2.0f * x);
    return 1.0f - __fdividef(2.0f, e + 1.0f);
}

// Build one [rows_pad x kc] K-major canonical tile from a row-major global matrix whose element (r, k)
// sits at flat index base + r*ld + k; `perturbed` applies theta + sg*eps, otherwise plain copy (observations).
// Elements outside [0,rows_real) x [0,k_real) are zero.
template <bool PERTURBED>
__device__ __forceinline__ void stage_tile(float* __restrict__ dst, const float* __restrict__ src0,
                                           const float* __restrict__ src1, float sg, int64_t base, int ld,
                                           int rows_real, int rows_pad, int k0, int k_real, int kc, int tid) {
    const int kq_n = kc >> 2;
    const bool aligned = ((base & 3) == 0) && ((ld & 3) == 0) && ((k0 & 3) == 0) &&
                         ((((uintptr_t)src0) & 15) == 0) && (!PERTURBED || ((((uintptr_t)src1) & 15) == 0));
    if (aligned) {
        const int items = (rows_pad >> 3) * kq_n * 8;
        for (int t = tid; t < items; t += TC_THREADS) {
            const int r8 = t & 7, q = t >> 3;
            const int kq = q % kq_n, rg = q / kq_n;
            const int r = rg * 8 + r8, k = k0 + kq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < rows_real && k < k_real) {
                const int64_t p = base + (int64_t)r * ld + k;
                if (k + 3 < k_real) {
                    const float4 a = *reinterpret_cast<const float4*>(src0 + p);
                    if (PERTURBED) {
                        const float4 e = ldg_stream_f4(src1 + p);
                        v = make_float4(perturb1(a.x, sg, e.x), perturb1(a.y, sg, e.y), perturb1(a.z, sg, e.z),
                                        perturb1(a.w, sg, e.w));
                    } else {
                        v = a;
                    }
                } else {
                    float t4[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int j = 0; j < 4 && k + j < k_real; ++j)
                        t4[j] = PERTURBED ? perturb1(src0[p + j], sg, src1[p + j]) : src0[p + j];
                    v = make_float4(t4[0], t4[1], t4[2], t4[3]);
                }
            }
            v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w);
            *reinterpret_cast<float4*>(dst + rg * (kc * 8) + kq * 32 + r8 * 4) = v;
        }
    } else {
        const int items = (rows_pad >> 3) * kq_n * 32;
        for (int t = tid; t < items; t += TC_THREADS) {
            const int r8 = t & 7, kk = (t >> 3) & 3, q = t >> 5;
            const int kq = q % kq_n, rg = q / kq_n;
            const int r = rg * 8 + r8, k = k0 + kq * 4 + kk;
            float v = 0.f;
            if (r < rows_real && k < k_real) {
                const int64_t p = base + (int64_t)r * ld + k;
                v = PERTURBED ? perturb1(src0[p], sg, src1[p]) : src0[p];
            }
            dst[rg * (kc * 8) + kq * 32 + r8 * 4 + kk] = to_tf32(v);
        }
    }
}

__global__ void __launch_bounds__(TC_THREADS) mlp_forward_tc_kernel(TcLayout L, const float* __restrict__ replicas,
                                                                    int64_t stride, const float* __restrict__ theta,
                                                                    const int64_t* __restrict__ idx,
                                                                    const int8_t* __restrict__ sign, float sigma,
                                                                    const float* __restrict__ obs, int E, int tiles,
                                                                    float* __restrict__ out, int hmax, int wbuf_floats,
                                                                    int tmem_cols) {
    extern __shared__ __align__(128) float smem[];
    // [ Hbuf: 128 x hmax (A0 chunks alias its head) | Wbuf | bias: N1 + N2 + N3 ]
    float* Hbuf = smem;
    float* Wbuf = Hbuf + 128 * hmax;
    float* bias = Wbuf + wbuf_floats;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m = blockIdx.x / tiles, tile = blockIdx.x % tiles;
    const int e0 = tile * 128;
    const int ne = min(128, E - e0);
    const float sg = sigma * (float)sign[m];
    const float* row = table_row_ptr(replicas, stride, idx[m]);
    const uint32_t bar = smem_u32(&mbar);

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // perturbed biases of all three layers (padded head entries are zero)
    for (int t = tid; t < L.N1 + L.N2 + L.N3; t += TC_THREADS) {
        int l = t < L.N1 ? 0 : (t < L.N1 + L.N2 ? 1 : 2);
        const int j = t - (l == 0 ? 0 : (l == 1 ? L.N1 : L.N1 + L.N2));
        float v = 0.f;
        if (j < L.nreal[l]) {
            const int p = L.b_off[l] + j;
            v = perturb1(theta[p], sg, row[p]);
        }
        bias[t] = v;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t d_col[3] = {0u, (uint32_t)L.N1, 0u};  // D3 reuses D1's columns (D1 is consumed by then)

    uint32_t phase = 0;
    const float* ob = obs + ((int64_t)m * E + e0) * L.K0;
    int bias_base = 0;
    for (int l = 0; l < 3; ++l) {
        const int Kp = L.kpad[l], N = L.npad[l];
        const uint32_t idesc = make_idesc_tf32(N);
        const uint32_t d_tmem = tmem + d_col[l];
        int kdone = 0;
        while (kdone < Kp) {
            const int kc = min(TC_KC, Kp - kdone);
            // --- producers: all threads build this chunk's operands in shared memory ---------------------
            stage_tile<true>(Wbuf, theta, row, sg, L.w_off[l], L.kin[l], L.nreal[l], N, kdone, L.kin[l], kc, tid);
            if (l == 0) stage_tile<false>(Hbuf, ob, nullptr, 0.f, 0, L.K0, ne, 128, kdone, L.K0, kc, tid);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // layer 0: A chunk is [128 x kc] on its own; layers 1,2: A is the full [128 x K] tile in Hbuf
                const int a_kc = (l == 0) ? kc : Kp;
                const uint32_t a_base = smem_u32(Hbuf) + (l == 0 ? 0u : (uint32_t)(kdone >> 2) * 128u);
                const uint32_t b_base = smem_u32(Wbuf);
                for (int j = 0; j < kc / 8; ++j) {
                    const uint64_t adesc = make_desc(a_base + j * 256, 128, (uint32_t)a_kc * 32u);
                    const uint64_t bdesc = make_desc(b_base + j * 256, 128, (uint32_t)kc * 32u);
                    umma_tf32(d_tmem, adesc, bdesc, idesc, (kdone > 0 || j > 0) ? 1u : 0u);
                }
                umma_commit(bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            kdone += kc;
        }
        // --- epilogue: TMEM -> registers -> bias + tanh -> next A operand (or the head) -------------------
        const int q = warp & 3;                 // TMEM lane quarter this warp may touch
        const int r = q * 32 + lane;            // observation row of this thread
        const int half = warp >> 2;             // column half
        const int ncol = (l < 2) ? N : L.N3;
        const int c_begin = half * (ncol / 2), c_end = c_begin + ncol / 2;
        if (l < 2) {
            const int Kn = L.kpad[l + 1];       // = N: width of the next layer's A tile
            for (int c = c_begin; c < c_end; c += 16) {
                float v[16];
                tmem_ld16(d_tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = to_tf32(tanh_fast(v[i] + bias[bias_base + c + i]));
                float* dst = Hbuf + (r >> 3) * (Kn * 8) + (c >> 2) * 32 + (r & 7) * 4;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<float4*>(dst + i * 32) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
        } else {
            float* o = out + ((int64_t)m * E + e0 + r) * L.nout;
            for (int c = c_begin; c < c_end; c += 16) {
                float v[16];
                tmem_ld16(d_tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
                if (r < ne) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int j = c + i;
                        if (j < L.nout) {
                            const float y = tanh_fast(v[i] + bias[bias_base + j]);
                            o[j] = j < L.A ? y : 0.55f + 0.45f * y;   // MapContinuousToAction
                        }
                    }
                }
            }
        }
        bias_base += N;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();   // H complete (and D consumed) before the next layer's producers / MMAs
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)tmem_cols) : "memory");
    }
}

}  // namespace

int dfd_mlp_forward_tc_impl(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                            const int64_t* idx, const int8_t* sign, int n_members, float sigma, const float* obs,
                            int obs_per_member, float* out, cudaStream_t st) {
    TcLayout L = {};
    L.K0 = desc->n_in;
    L.K0p = (desc->n_in + 7) / 8 * 8;
    L.N1 = desc->h1;
    L.N2 = desc->h2;
    L.A = desc->n_act;
    L.nout = 2 * desc->n_act;
    L.N3 = (L.nout + 31) / 32 * 32;   // epilogue splits the head columns in two halves of 16-column loads
    DFD_CHECK_ARG(L.N1 % 32 == 0 && L.N2 % 32 == 0 && L.N1 <= 256 && L.N2 <= 256,
                  "tcgen05 MLP path: hidden widths must be multiples of 32 and <= 256 (got %d, %d)", L.N1, L.N2);
    DFD_CHECK_ARG(L.N3 <= 256 && L.N3 <= L.N1, "tcgen05 MLP path: head width %d too large", L.nout);
    const int in_[3] = {L.K0, L.N1, L.N2}, outr[3] = {L.N1, L.N2, L.nout}, outp[3] = {L.N1, L.N2, L.N3};
    int off = 0;
    for (int l = 0; l < 3; ++l) {
        L.w_off[l] = off; off += in_[l] * outr[l];
        L.b_off[l] = off; off += outr[l];
        L.kin[l] = in_[l];
        L.kpad[l] = l == 0 ? L.K0p : in_[l];
        L.nreal[l] = outr[l];
        L.npad[l] = outp[l];
    }
    L.P = off;
    const int hmax = L.N1 > L.N2 ? L.N1 : L.N2;
    int wbuf = 0;
    for (int l = 0; l < 3; ++l) {
        const int kc = L.kpad[l] < TC_KC ? L.kpad[l] : TC_KC;
        if (outp[l] * kc > wbuf) wbuf = outp[l] * kc;
    }
    DFD_CHECK_ARG(128 * (L.K0p < TC_KC ? L.K0p : TC_KC) <= 128 * hmax, "tcgen05 MLP path: observation chunk exceeds H buffer");
    int cols = 32;
    while (cols < L.N1 + L.N2) cols <<= 1;
    DFD_CHECK_ARG(cols <= 512, "tcgen05 MLP path: needs %d TMEM columns", cols);
    const size_t smem = ((size_t)128 * hmax + wbuf + L.N1 + L.N2 + L.N3) * sizeof(float) + 128;
    DFD_CHECK_ARG(smem <= 227 * 1024, "tcgen05 MLP path: needs %zu B shared memory", smem);
    const int tiles = (obs_per_member + 127) / 128;
    DFD_CHECK_ARG((int64_t)n_members * tiles < 2147483647LL, "tcgen05 MLP path: grid too large");
    DFD_CUDA(cudaFuncSetAttribute(mlp_forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp_forward_tc_kernel<<<n_members * tiles, TC_THREADS, smem, st>>>(L, table->replicas, table->replica_stride, theta,
                                                                       idx, sign, sigma, obs, obs_per_member, tiles, out,
                                                                       hmax, wbuf, cols);
    DFD_LAUNCHED(ctx);
    return 0;
}
