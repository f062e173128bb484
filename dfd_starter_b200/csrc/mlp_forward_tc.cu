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
#include "tc_common.cuh"

namespace {

// Build one [rows_pad x kc] K-major canonical tile from a row-major global matrix whose element (r, k)
// sits at flat index base + r*ld + k; PERTURBED applies theta + sg*eps, otherwise plain copy (observations).
// Elements outside [0,rows_real) x [k0, k_real) are zero.  Thread -> element mapping is fixed (no divisions):
// 8 consecutive lanes take the 8 rows of a core matrix so every quarter-warp writes one contiguous 128-byte
// core-matrix row block (conflict-free), and a warp reads 64 contiguous bytes from each of 8 rows.
template <bool PERTURBED>
__device__ __forceinline__ void stage_tile(float* __restrict__ dst, const float* __restrict__ src0,
                                           const float* __restrict__ src1, float sg, int64_t base, int ld,
                                           int rows_real, int rows_pad, int k0, int k_real, int kc, int tid) {
    const int kq_n = kc >> 2, rg_n = rows_pad >> 3;
    const int r8 = tid & 7;
    const bool aligned = ((base & 3) == 0) && ((ld & 3) == 0) && ((k0 & 3) == 0) &&
                         ((((uintptr_t)src0) & 15) == 0) && (!PERTURBED || ((((uintptr_t)src1) & 15) == 0));
    if (aligned) {
        // 16-byte items: lane -> (r8, kq = (tid>>3)&15, rg = tid>>7); a pass covers 2 row groups x 64 columns
        const int kql = (tid >> 3) & 15, rgl = tid >> 7;
        constexpr int B = 4;
        for (int kq0 = 0; kq0 < kq_n; kq0 += 16) {
            const int kq = kq0 + kql;
            const int k = k0 + kq * 4;
            const bool k_ok = kq < kq_n;
            for (int rg0 = 0; rg0 < rg_n; rg0 += 2 * B) {
                float4 a[B], e[B];
#pragma unroll
                for (int b = 0; b < B; ++b) {
                    const int rg = rg0 + 2 * b + rgl, r = rg * 8 + r8;
                    a[b] = make_float4(0.f, 0.f, 0.f, 0.f);
                    e[b] = a[b];
                    if (k_ok && rg < rg_n && r < rows_real && k < k_real) {
                        const int64_t p = base + (int64_t)r * ld + k;
                        if (k + 3 < k_real) {
                            a[b] = *reinterpret_cast<const float4*>(src0 + p);
                            if (PERTURBED) e[b] = ldg_stream_f4(src1 + p);
                        } else {   // ragged K edge (only when k_real is not a multiple of 4)
                            float t4[4] = {0.f, 0.f, 0.f, 0.f}, e4[4] = {0.f, 0.f, 0.f, 0.f};
                            for (int j = 0; j < 4 && k + j < k_real; ++j) {
                                t4[j] = src0[p + j];
                                if (PERTURBED) e4[j] = src1[p + j];
                            }
                            a[b] = make_float4(t4[0], t4[1], t4[2], t4[3]);
                            e[b] = make_float4(e4[0], e4[1], e4[2], e4[3]);
                        }
                    }
                }
#pragma unroll
                for (int b = 0; b < B; ++b) {
                    const int rg = rg0 + 2 * b + rgl;
                    if (k_ok && rg < rg_n) {
                        float4 v = a[b];
                        if (PERTURBED)
                            v = make_float4(perturb1(a[b].x, sg, e[b].x), perturb1(a[b].y, sg, e[b].y),
                                            perturb1(a[b].z, sg, e[b].z), perturb1(a[b].w, sg, e[b].w));
                        v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w);
                        *reinterpret_cast<float4*>(dst + rg * (kc * 8) + kq * 32 + r8 * 4) = v;
                    }
                }
            }
        }
    } else {
        // 4-byte items: lane -> (r8, kk = (tid>>3)&3, kq = tid>>5); a pass covers 1 row group x 32 columns
        const int kk = (tid >> 3) & 3, kql = tid >> 5;
        constexpr int B = 8;
        for (int kq0 = 0; kq0 < kq_n; kq0 += 8) {
            const int kq = kq0 + kql;
            const int k = k0 + kq * 4 + kk;
            const bool k_ok = kq < kq_n;
            for (int rg0 = 0; rg0 < rg_n; rg0 += B) {
                float a[B], e[B];
#pragma unroll
                for (int b = 0; b < B; ++b) {
                    const int rg = rg0 + b, r = rg * 8 + r8;
                    a[b] = 0.f; e[b] = 0.f;
                    if (k_ok && rg < rg_n && r < rows_real && k < k_real) {
                        const int64_t p = base + (int64_t)r * ld + k;
                        a[b] = src0[p];
                        if (PERTURBED) e[b] = src1[p];
                    }
                }
#pragma unroll
                for (int b = 0; b < B; ++b) {
                    const int rg = rg0 + b;
                    if (k_ok && rg < rg_n)
                        dst[rg * (kc * 8) + kq * 32 + r8 * 4 + kk] = to_tf32(PERTURBED ? perturb1(a[b], sg, e[b]) : a[b]);
                }
            }
        }
    }
}

// resident = 1: all three canonical weight tiles of a member fit in shared memory next to the activation
// tile (the 64x64 nets): everything a member needs is staged in ONE pass (one exposed memory latency), then
// the three MMA + epilogue rounds run from shared memory / TMEM only.  resident = 0: weights stream through
// one buffer in K chunks of TC_KC columns (the 256x256 Humanoid net).
__global__ void __launch_bounds__(TC_THREADS, 3) mlp_forward_tc_kernel(TcLayout L, const float* __restrict__ replicas,
                                                                    int64_t stride, const float* __restrict__ theta,
                                                                    const int64_t* __restrict__ idx,
                                                                    const int8_t* __restrict__ sign, float sigma,
                                                                    const float* __restrict__ obs, int E, int tiles,
                                                                    int n_work, float* __restrict__ out, int hmax,
                                                                    int resident, int w_off0, int w_off1, int w_off2,
                                                                    int wbuf_floats, int tmem_cols, long long* __restrict__ prof) {
    extern __shared__ __align__(128) float smem[];
    // [ Hbuf: 128 x hmax (the observation tile / chunks alias its head) | Wbuf | bias: N1 + N2 + N3 ]
    float* Hbuf = smem;
    float* Wbuf = Hbuf + 128 * hmax;
    float* bias = Wbuf + wbuf_floats;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar = smem_u32(&mbar);
    const int wofs[3] = {w_off0, w_off1, w_off2};

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
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t d_col[3] = {0u, (uint32_t)L.N1, 0u};  // D3 reuses D1's columns (D1 is consumed by then)
    uint32_t phase = 0;

#define TC_STAMP(i) do { if (prof && tid == 0 && work == (int)blockIdx.x + (int)gridDim.x) prof[blockIdx.x * 32 + (i)] = clock64(); } while (0)
    for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
        TC_STAMP(0);
        const int m = work / tiles, tile = work - m * tiles;
        const int e0 = tile * 128;
        const int ne = min(128, E - e0);
        const float sg = sigma * (float)sign[m];
        const float* row = table_row_ptr(replicas, stride, idx[m]);
        const float* ob = obs + ((int64_t)m * E + e0) * L.K0;
        if (tid == 0) {
            const int nw = work + gridDim.x;
            if (nw < n_work) {
                const int nm = nw / tiles, nt = nw - nm * tiles;
                if (nt == 0 || tiles == 1) l2_prefetch(table_row_ptr(replicas, stride, idx[nm]), (size_t)L.P * 4);
                l2_prefetch(obs + ((int64_t)nm * E + nt * 128) * L.K0, (size_t)min(128, E - nt * 128) * L.K0 * 4);
            }
        }

        // perturbed biases of all three layers (padded head entries are zero)
        for (int t = tid; t < L.N1 + L.N2 + L.N3; t += TC_THREADS) {
            const int l = t < L.N1 ? 0 : (t < L.N1 + L.N2 ? 1 : 2);
            const int j = t - (l == 0 ? 0 : (l == 1 ? L.N1 : L.N1 + L.N2));
            float v = 0.f;
            if (j < L.nreal[l]) {
                const int p = L.b_off[l] + j;
                v = perturb1(theta[p], sg, row[p]);
            }
            bias[t] = v;
        }
        TC_STAMP(1);
        if (resident == 2) {
            // small nets (widths <= 64, K0p <= 32): two load bursts per member instead of one load->use round
            // trip per tile.  Burst A: W1, W2 as 16-byte items; burst B: W0 and the observation tile as 4-byte
            // items (rows of 17 floats are not 16-byte aligned).  Same lane -> element maps as stage_tile.
            const int r8 = tid & 7;
            {
                const int kq = (tid >> 3) & 15, rgl = tid >> 7;
                float4 a1[4], e1[4], a2[2], e2[2];
                const int kq1 = L.kpad[1] >> 2, rg1 = L.npad[1] >> 3, kq2 = L.kpad[2] >> 2, rg2 = L.npad[2] >> 3;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int rg = 2 * b + rgl, rr = rg * 8 + r8;
                    a1[b] = make_float4(0.f, 0.f, 0.f, 0.f); e1[b] = a1[b];
                    if (kq < kq1 && rg < rg1 && rr < L.nreal[1]) {
                        const int64_t p = L.w_off[1] + (int64_t)rr * L.kin[1] + kq * 4;
                        a1[b] = *reinterpret_cast<const float4*>(theta + p);
                        e1[b] = ldg_stream_f4(row + p);
                    }
                }
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int rg = 2 * b + rgl, rr = rg * 8 + r8;
                    a2[b] = make_float4(0.f, 0.f, 0.f, 0.f); e2[b] = a2[b];
                    if (kq < kq2 && rg < rg2 && rr < L.nreal[2]) {
                        const int64_t p = L.w_off[2] + (int64_t)rr * L.kin[2] + kq * 4;
                        a2[b] = *reinterpret_cast<const float4*>(theta + p);
                        e2[b] = ldg_stream_f4(row + p);
                    }
                }
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int rg = 2 * b + rgl;
                    if (kq < kq1 && rg < rg1) {
                        float4 v = make_float4(perturb1(a1[b].x, sg, e1[b].x), perturb1(a1[b].y, sg, e1[b].y),
                                               perturb1(a1[b].z, sg, e1[b].z), perturb1(a1[b].w, sg, e1[b].w));
                        v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w);
                        *reinterpret_cast<float4*>(Wbuf + wofs[1] + rg * (L.kpad[1] * 8) + kq * 32 + r8 * 4) = v;
                    }
                }
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int rg = 2 * b + rgl;
                    if (kq < kq2 && rg < rg2) {
                        float4 v = make_float4(perturb1(a2[b].x, sg, e2[b].x), perturb1(a2[b].y, sg, e2[b].y),
                                               perturb1(a2[b].z, sg, e2[b].z), perturb1(a2[b].w, sg, e2[b].w));
                        v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w);
                        *reinterpret_cast<float4*>(Wbuf + wofs[2] + rg * (L.kpad[2] * 8) + kq * 32 + r8 * 4) = v;
                    }
                }
            }
            {
                const int kk = (tid >> 3) & 3, kq = tid >> 5, k = kq * 4 + kk;
                const int kq0n = L.K0p >> 2, rg0n = L.npad[0] >> 3;
                float a0[8], e0[8], ao[16];
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int rr = b * 8 + r8;
                    a0[b] = 0.f; e0[b] = 0.f;
                    if (kq < kq0n && b < rg0n && rr < L.nreal[0] && k < L.K0) {
                        const int64_t p = L.w_off[0] + (int64_t)rr * L.K0 + k;
                        a0[b] = theta[p];
                        e0[b] = row[p];
                    }
                }
#pragma unroll
                for (int b = 0; b < 16; ++b) {
                    const int rr = b * 8 + r8;
                    ao[b] = (kq < kq0n && rr < ne && k < L.K0) ? ob[(int64_t)rr * L.K0 + k] : 0.f;
                }
#pragma unroll
                for (int b = 0; b < 8; ++b)
                    if (kq < kq0n && b < rg0n)
                        Wbuf[wofs[0] + b * (L.K0p * 8) + kq * 32 + r8 * 4 + kk] = to_tf32(perturb1(a0[b], sg, e0[b]));
#pragma unroll
                for (int b = 0; b < 16; ++b)
                    if (kq < kq0n) Hbuf[b * (L.K0p * 8) + kq * 32 + r8 * 4 + kk] = to_tf32(ao[b]);
            }
        } else if (resident) {
            for (int l = 0; l < 3; ++l)
                stage_tile<true>(Wbuf + wofs[l], theta, row, sg, L.w_off[l], L.kin[l], L.nreal[l], L.npad[l], 0, L.kin[l],
                                 L.kpad[l], tid);
            stage_tile<false>(Hbuf, ob, nullptr, 0.f, 0, L.K0, ne, 128, 0, L.K0, L.K0p, tid);
        }
        TC_STAMP(2);
        int bias_base = 0;
        for (int l = 0; l < 3; ++l) {
            const int Kp = L.kpad[l], N = L.npad[l];
            const uint32_t idesc = make_idesc_tf32(N);
            const uint32_t d_tmem = tmem + d_col[l];
            int kdone = 0;
            while (kdone < Kp) {
                const int kc = resident ? Kp : min(TC_KC, Kp - kdone);
                if (!resident) {
                    stage_tile<true>(Wbuf, theta, row, sg, L.w_off[l], L.kin[l], L.nreal[l], N, kdone, L.kin[l], kc, tid);
                    if (l == 0) stage_tile<false>(Hbuf, ob, nullptr, 0.f, 0, L.K0, ne, 128, kdone, L.K0, kc, tid);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
                __syncthreads();
                TC_STAMP(3 + 4 * l);
                if (warp == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    // layer 0, streamed: the A chunk is [128 x kc] on its own; otherwise A is a full [128 x Kp] tile
                    const int a_kc = (l == 0 && !resident) ? kc : Kp;
                    const uint32_t a_base = smem_u32(Hbuf) + ((l == 0 && !resident) ? 0u : (uint32_t)(kdone >> 2) * 128u);
                    const uint32_t b_base = smem_u32(Wbuf + (resident ? wofs[l] : 0));
                    const uint64_t adesc = make_desc(a_base, 128, (uint32_t)a_kc * 32u);
                    const uint64_t bdesc = make_desc(b_base, 128, (uint32_t)kc * 32u);
                    umma_tf32_elect(d_tmem, adesc, bdesc, idesc, kdone > 0 ? 1u : 0u);
#pragma unroll 4
                    for (int j = 1; j < kc / 8; ++j)
                        umma_tf32_elect(d_tmem, adesc + (uint64_t)(j * 16), bdesc + (uint64_t)(j * 16), idesc, 1u);
                    umma_commit_elect(bar);
                    __syncwarp();
                }
                TC_STAMP(4 + 4 * l);
                if (warp == 0) mbar_wait(bar, phase);   // one warp polls; the rest sleep at the barrier
                phase ^= 1;
                __syncthreads();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                kdone += kc;
                TC_STAMP(5 + 4 * l);
            }
            // --- epilogue: TMEM -> registers -> bias + tanh -> next A operand (or the head) ---------------
            const int q = warp & 3;                 // TMEM lane quarter this warp may touch
            const int r = q * 32 + lane;            // observation row of this thread
            const int half = warp >> 2;             // column half
            const int ncol = (l < 2) ? N : L.N3;
            const int c_begin = half * (ncol / 2), c_end = c_begin + ncol / 2;
            if (l < 2) {
                const int Kn = L.kpad[l + 1];       // = N: width of the next layer's A tile
                for (int c = c_begin; c < c_end; c += 32) {   // hidden widths are multiples of 64 halves of 32
                    float v[32];
                    tmem_ld32(d_tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
                    const float4* b4 = reinterpret_cast<const float4*>(bias + bias_base + c);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 bb = b4[i];
                        v[4 * i + 0] = to_tf32(tanh_fast(v[4 * i + 0] + bb.x));
                        v[4 * i + 1] = to_tf32(tanh_fast(v[4 * i + 1] + bb.y));
                        v[4 * i + 2] = to_tf32(tanh_fast(v[4 * i + 2] + bb.z));
                        v[4 * i + 3] = to_tf32(tanh_fast(v[4 * i + 3] + bb.w));
                    }
                    float* dst = Hbuf + (r >> 3) * (Kn * 8) + (c >> 2) * 32 + (r & 7) * 4;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        *reinterpret_cast<float4*>(dst + i * 32) =
                            make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
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
            TC_STAMP(6 + 4 * l);
        }
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
    DFD_CHECK_ARG(L.N1 % 64 == 0 && L.N2 % 64 == 0 && L.N1 <= 256 && L.N2 <= 256,
                  "tcgen05 MLP path: hidden widths must be multiples of 64 and <= 256 (got %d, %d)", L.N1, L.N2);
    DFD_CHECK_ARG(L.N3 <= 256 && L.N3 <= L.N1 && (L.N1 + L.N2 + L.N3) % 4 == 0, "tcgen05 MLP path: head width %d too large", L.nout);
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
    int hmax = L.N1 > L.N2 ? L.N1 : L.N2;
    if (L.K0p > hmax && L.K0p <= 256) hmax = L.K0p;   // the resident observation tile lives in the H buffer
    // resident mode: all three canonical weight tiles + the activation tile fit comfortably (>= 2 CTAs / SM)
    const int wtile[3] = {outp[0] * L.kpad[0], outp[1] * L.kpad[1], outp[2] * L.kpad[2]};
    const size_t res_bytes = ((size_t)128 * hmax + wtile[0] + wtile[1] + wtile[2] + L.N1 + L.N2 + L.N3) * sizeof(float) + 128;
    int resident = (res_bytes <= 100 * 1024 && L.K0p <= hmax) ? 1 : 0;
    // burst staging needs: widths <= 64, head <= 32 padded rows, K0p <= 32, 16-byte aligned W1 / W2 rows
    if (resident && L.N1 <= 64 && L.N2 <= 64 && L.N3 <= 32 && L.K0p <= 32 && (L.w_off[1] & 3) == 0 &&
        (L.w_off[2] & 3) == 0 && (L.kin[1] & 3) == 0 && (L.kin[2] & 3) == 0 && (((uintptr_t)theta) & 15) == 0)
        resident = 2;
    if (const char* e = getenv("DFD_TC_MODE")) {   // experiment switch: 1 = per-tile staging, 2 = burst staging
        if (resident && atoi(e) >= 1 && atoi(e) <= resident) resident = atoi(e);
    }
    int wbuf = 0;
    if (resident) {
        wbuf = wtile[0] + wtile[1] + wtile[2];
    } else {
        hmax = L.N1 > L.N2 ? L.N1 : L.N2;
        for (int l = 0; l < 3; ++l) {
            const int kc = L.kpad[l] < TC_KC ? L.kpad[l] : TC_KC;
            if (outp[l] * kc > wbuf) wbuf = outp[l] * kc;
        }
        DFD_CHECK_ARG(128 * (L.K0p < TC_KC ? L.K0p : TC_KC) <= 128 * hmax, "tcgen05 MLP path: observation chunk exceeds H buffer");
    }
    int cols = 32;
    while (cols < L.N1 + L.N2) cols <<= 1;
    DFD_CHECK_ARG(cols <= 512, "tcgen05 MLP path: needs %d TMEM columns", cols);
    const size_t smem = ((size_t)128 * hmax + wbuf + L.N1 + L.N2 + L.N3) * sizeof(float) + 128;
    DFD_CHECK_ARG(smem <= 227 * 1024, "tcgen05 MLP path: needs %zu B shared memory", smem);
    const int tiles = (obs_per_member + 127) / 128;
    DFD_CHECK_ARG((int64_t)n_members * tiles < 2147483647LL, "tcgen05 MLP path: too many work items");
    const int n_work = n_members * tiles;
    DFD_CUDA(cudaFuncSetAttribute(mlp_forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // persistent CTAs: as many as fit per SM (shared memory, TMEM columns, 8 warps each), looping over work items
    // resident CTAs per SM: shared memory, registers and TMEM columns (the occupancy API does not know TMEM)
    DFD_CUDA(cudaFuncSetAttribute(mlp_forward_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  (int)cudaSharedmemCarveoutMaxShared));
    cudaFuncAttributes fa;
    DFD_CUDA(cudaFuncGetAttributes(&fa, mlp_forward_tc_kernel));
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    const int by_regs = 65536 / ((fa.numRegs + 7) / 8 * 8 * TC_THREADS);
    if (per_sm > by_regs) per_sm = by_regs;
    if (per_sm > 512 / cols) per_sm = 512 / cols;
    if (per_sm < 1) per_sm = 1;
    int grid = ctx->sm_count * per_sm;
    if (grid > n_work) grid = n_work;
    long long* prof = nullptr;
    static const bool want_prof = getenv("DFD_TC_PROF") != nullptr;   // debugging aid: per-phase clock64 stamps
    if (want_prof) {
        cudaMalloc(&prof, (size_t)grid * 32 * sizeof(long long));
        cudaMemset(prof, 0, (size_t)grid * 32 * sizeof(long long));
    }
    mlp_forward_tc_kernel<<<grid, TC_THREADS, smem, st>>>(L, table->replicas, table->replica_stride, theta, idx, sign,
                                                          sigma, obs, obs_per_member, tiles, n_work, out, hmax, resident,
                                                          0, wtile[0], wtile[0] + wtile[1], wbuf, cols, prof);
    DFD_LAUNCHED(ctx);
    if (want_prof) {
        cudaStreamSynchronize(st);
        static long long h[1024 * 32];
        const int g = grid < 1024 ? grid : 1024;
        cudaMemcpy(h, prof, (size_t)g * 32 * sizeof(long long), cudaMemcpyDeviceToHost);
        double acc[16] = {0};
        int cnt = 0;
        for (int b = 0; b < g; ++b) {
            if (h[b * 32 + 0] == 0 || h[b * 32 + 14] == 0) continue;
            for (int i = 1; i < 15; ++i) acc[i] += (double)(h[b * 32 + i] - h[b * 32 + i - 1]);
            ++cnt;
        }
        if (cnt) {
            fprintf(stderr, "[tc prof] grid %d resident %d, mean cycles per phase over %d CTAs (2nd work item):\n", grid, resident, cnt);
            const char* nm[15] = {"", "bias+prefetch", "stage all", "fence+sync L0", "mma issue L0", "mma wait L0", "epilogue L0",
                                  "fence+sync L1", "mma issue L1", "mma wait L1", "epilogue L1", "fence+sync L2", "mma issue L2",
                                  "mma wait L2", "epilogue L2"};
            for (int i = 1; i < 15; ++i) fprintf(stderr, "   %-16s %8.0f\n", nm[i], acc[i] / cnt);
        }
        cudaFree(prof);
    }
    return 0;
}
