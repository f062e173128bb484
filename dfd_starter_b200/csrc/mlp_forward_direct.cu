// "Direct-from-table" per-member MLP forward for the wide MuJoCo nets (Humanoid-shaped 376-256-256-17 of BASELINE config 3;
// policies/mujoco.py:35-41 with the perturbation of worker/worker.py:28).
//
// A member's weights are theta + s*sigma*eps and a Linear layer is linear in its weights:
//     x . (theta + s*sigma*eps)^T  =  x . theta^T  +  s * (x . (sigma*eps)^T)
// so the perturbed weights are NEVER BUILT.  Both terms are tcgen05 MMAs into the same TMEM accumulator whose B operands
// arrive by TMA straight from global memory in the UMMA operand layout (SWIZZLE_128B, K-major):
//   * W tiles from an fp16 copy of theta (made once per call by theta_to_f16_kernel; every member reads the same tiles,
//     they stay in L2);
//   * E tiles from a sigma-scaled fp16 mirror of the noise table (dfd_table_build_scaled16: eight element-shifted
//     replicas, so every table[idx + off ...] slice starts 16-byte aligned in replica (idx + off) & 7), addressed through
//     a 4-D tensor map whose second dimension OVERLAPS the first - dims {K, start / 8, rows, replica}, strides {16 B,
//     K * 2 B, replica bytes} - which makes the row-major [N x K] weight matrix of ANY member a legal TMA box;
//   * the sign of the member is the negate-A bit of the instruction descriptor of the E MMAs.
// No thread touches a weight.  fp16 operands (10-bit mantissa, the precision of tf32), fp32 accumulate; biases are
// perturbed exactly in fp32 (theta_b + s*sigma*eps_b, two roundings) and added in the epilogue.
//
//   warp 0        TMA producer: W / E tiles [128 rows x 64 k] (16 KB) into a ring of shared-memory slots
//   warp 1        MMA issue: kind::f16, M = 128 observations, N = 128 (head: 48); layer 0 takes A from shared memory,
//                 layers 1 / 2 take A from TENSOR MEMORY (packed halves)
//   warp 2        perturbed biases of the next member into shared memory; TMEM allocation
//   warps 4-11    epilogue, two warps per TMEM lane quarter: accumulator -> + bias -> tanh -> fp16x2 -> the next layer's A
//                 operand in TMEM, 32 activations at a time, each batch published on its own mbarrier
//   warps 12-15   observation tile fp32 -> fp16 into the swizzled A-operand slots of layer 0
// TMEM (512 columns):  [0,256) layer-0 accumulator, later [0,128) layer-1 accumulator (outputs 128..255) and [128,256) the
// layer-2 A operand;  [256,384) layer-1 A operand;  [384,512) layer-1 accumulator (outputs 0..127), later the head accumulator.
#include "direct_common.cuh"

namespace {

constexpr int DR_KC = 64;                  // halves per tile row: one 128-byte swizzle row
constexpr int DR_TILE = 128 * DR_KC * 2;   // bytes of a [128 x 64] fp16 tile
constexpr int DR_NS_MAX = 10, DR_NX_MAX = 3;
constexpr int DR_THREADS = 512;
constexpr int DR_EPI_WARP0 = 4, DR_CVT_WARP0 = 12;

struct DrParams {
    int K0, N1, N2, nout, N3, A;
    int w_off[3], b_off[3], kin[3], nreal[3];
    int E, tiles, n_work, pair_order, ns, nx;
    int64_t P;
    float sigma;
};

struct DrMaps {
    CUtensorMap w[3];     // fp16 theta copy: 2-D [n_out rows, k_in], box {64, 128} (head: {64, N3})
    CUtensorMap e[3];     // scaled fp16 table: 4-D {k_in, starts, n_out, 8 replicas}, box {64, 1, 128, 1} (head: rows N3)
};

enum {
    DB_FULL = 0, DB_EMPTY = DB_FULL + DR_NS_MAX, DB_XFULL = DB_EMPTY + DR_NS_MAX, DB_XEMPTY = DB_XFULL + DR_NX_MAX,
    DB_D0FULL = DB_XEMPTY + DR_NX_MAX, DB_D1FULL /* 2 */, DB_DHFULL = DB_D1FULL + 2, DB_A1READY /* 8 */, DB_A2READY = DB_A1READY + 8,
    DB_BFULL = DB_A2READY + 8 /* 2 */, DB_BEMPTY = DB_BFULL + 2 /* 2 */, DB_COUNT = DB_BEMPTY + 2
};

__device__ __forceinline__ void dr_item(const DrParams& p, int work, int& m, int& tile) {
    const int mm = p.tiles == 1 ? work : work / p.tiles;
    tile = work - mm * p.tiles;
    const int M = p.n_work / p.tiles;
    m = p.pair_order ? ((mm & 1) ? (M >> 1) + (mm >> 1) : (mm >> 1)) : mm;
}
// replica_s[j] = fp16(fl32(sigma * table[j + s])), s = 0..7: the first rounding is the reference's own (worker.py:28 rounds
// sigma*eps to fp32 before the add), the second is the operand precision of this path
__global__ void table_scaled16_kernel(const float* __restrict__ replica0, int64_t size, float sigma, __half* __restrict__ out,
                                      int64_t stride16) {
    const int64_t total = 8 * stride16;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = i / stride16, j = i - s * stride16;
        const int64_t src = j + s;
        out[i] = src < size ? __float2half_rn(__fmul_rn(sigma, replica0[src])) : __float2half_rn(0.f);
    }
}

template <bool APPROX>
__global__ void __launch_bounds__(DR_THREADS, 1)
mlp_forward_direct_kernel(const DrParams p, const __grid_constant__ DrMaps maps, const float* __restrict__ replicas, int64_t stride,
                          const float* __restrict__ theta, const int64_t* __restrict__ idx, const int8_t* __restrict__ sign,
                          const float* __restrict__ obs, float* __restrict__ out, long long* __restrict__ prof) {
#define DR_TL(cond, slot) do { if (prof && (cond) && u == 3) prof[(size_t)blockIdx.x * 32 + (slot)] = clock64(); } while (0)
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[DB_COUNT];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int n_my = ((int)blockIdx.x < p.n_work) ? (p.n_work - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const uint32_t bar0 = smem_u32(&bars[0]);
#define DR_BAR(i) (bar0 + 8u * (uint32_t)(i))
    const uint32_t smem_base = smem_u32(smem_raw);
    const uint32_t ring0 = (smem_base + 1023u) & ~1023u;                    // B-tile ring: ns x 16 KB (1024-byte aligned atoms)
    const uint32_t xring0 = ring0 + (uint32_t)p.ns * DR_TILE;                // A-tile ring of layer 0: nx x 16 KB
    float* bias_s = reinterpret_cast<float*>(smem_raw + (xring0 - smem_base) + (uint32_t)p.nx * DR_TILE);   // [2][3][256]
    float* ostage = bias_s + 2 * 768;                                        // [128 x nout]
    const int nh1 = p.N1 >> 7, nh2 = p.N2 >> 7;                              // 128-wide halves of the hidden layers
    const int nc0 = (p.K0 + DR_KC - 1) / DR_KC, nc1 = p.N1 / DR_KC, nc2 = p.N2 / DR_KC;

    if (tid == 0) {
        for (int s = 0; s < DR_NS_MAX; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(DR_BAR(DB_FULL + s)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(DR_BAR(DB_EMPTY + s)));
        }
        for (int s = 0; s < DR_NX_MAX; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(DR_BAR(DB_XFULL + s)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(DR_BAR(DB_XEMPTY + s)));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(DR_BAR(DB_D0FULL)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(DR_BAR(DB_D1FULL)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(DR_BAR(DB_D1FULL + 1)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(DR_BAR(DB_DHFULL)));
        for (int j = 0; j < 16; ++j) asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(DR_BAR(DB_A1READY + j)));
        for (int b = 0; b < 2; ++b) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(DR_BAR(DB_BFULL + b)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(DR_BAR(DB_BEMPTY + b)));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t tD0 = tmem, tA2 = tmem + 128u, tA1 = tmem + 256u, tD1a = tmem + 384u, tD1b = tmem, tDH = tmem + 384u;

    if (warp == 0) {
        // =============================== TMA producer ===========================================
        if (lane == 0) {
            int g = 0;
            for (int u = 0; u < n_my; ++u) {
                const int work = (int)blockIdx.x + u * (int)gridDim.x;
                int m, tile;
                dr_item(p, work, m, tile);
                const int sg = (int)sign[m];
                const int64_t id = idx[m];
#pragma unroll 1
                for (int l = 0; l < 3; ++l) {
                    const int64_t s = id + p.w_off[l];
                    const int st8 = (int)(s >> 3), rep = (int)(s & 7);
                    const int nc = l == 0 ? nc0 : (l == 1 ? nc1 : nc2);
                    const int nh = l == 0 ? nh1 : (l == 1 ? nh2 : 1);
                    const uint32_t bytes = l == 2 ? (uint32_t)p.N3 * 128u : (uint32_t)DR_TILE;
                    // layer 0: tiles in (chunk, half) order; layer 1: (half, chunk) - a half is a complete K loop into its own
                    // accumulator; head: chunk order
                    const int outer = l == 1 ? nh : nc, inner = l == 1 ? nc : nh;
#pragma unroll 1
                    for (int a = 0; a < outer; ++a)
#pragma unroll 1
                        for (int b = 0; b < inner; ++b) {
                            const int c = l == 1 ? b : a, h = l == 1 ? a : b;
#pragma unroll 1
                            for (int we = 0; we < 2; ++we) {
                                if (we == 1 && sg == 0) continue;           // unperturbed member: no E term
                                const int slot = g % p.ns;
                                dr_wait(DR_BAR(DB_EMPTY + slot), (uint32_t)((g / p.ns) & 1) ^ 1u);
                                DR_TL(l == 0 && a == 0 && b == 0 && we == 0, 16);
                                DR_TL(l == 1 && a == 0 && b == 0 && we == 0, 17);
                                DR_TL(l == 2 && a == 0 && we == 0, 18);
                                dr_expect_tx(DR_BAR(DB_FULL + slot), bytes);
                                const uint32_t dst = ring0 + (uint32_t)slot * DR_TILE;
                                if (we == 0) dr_tma_2d(dst, &maps.w[l], c * DR_KC, h * 128, DR_BAR(DB_FULL + slot));
                                else dr_tma_4d(dst, &maps.e[l], c * DR_KC, st8, h * 128, rep, DR_BAR(DB_FULL + slot));
                                ++g;
                            }
                        }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issue ==============================================
        int g = 0, xg = 0;
        uint32_t ph = 0;
        for (int u = 0; u < n_my; ++u) {
            const int work = (int)blockIdx.x + u * (int)gridDim.x;
            int m, tile;
            dr_item(p, work, m, tile);
            const int sg = (int)sign[m];
            const uint32_t id_w128 = dr_idesc(128, 0), id_e128 = dr_idesc(128, sg < 0);
            // ---- layer 0: A = observation tile in shared memory ----
            DR_TL(lane == 0, 0);
#pragma unroll 1
            for (int c = 0; c < nc0; ++c, ++xg) {
                const int xs = xg % p.nx;
                dr_wait(DR_BAR(DB_XFULL + xs), (uint32_t)((xg / p.nx) & 1));
                DR_TL(lane == 0 && c == 0, 7);
                const uint64_t adesc = make_desc_sw128(xring0 + (uint32_t)xs * DR_TILE);
#pragma unroll 1
                for (int h = 0; h < nh1; ++h)
#pragma unroll 1
                    for (int we = 0; we < 2; ++we) {
                        if (we == 1 && sg == 0) continue;
                        const int slot = g % p.ns;
                        dr_wait(DR_BAR(DB_FULL + slot), (uint32_t)((g / p.ns) & 1));
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t bdesc = make_desc_sw128(ring0 + (uint32_t)slot * DR_TILE);
#pragma unroll
                        for (int j = 0; j < DR_KC / 16; ++j)
                            dr_umma_ss(tD0 + (uint32_t)(h * 128), adesc + (uint64_t)(j * 2), bdesc + (uint64_t)(j * 2),
                                       we ? id_e128 : id_w128, (c | we | j) ? 1u : 0u);
                        umma_commit_elect(DR_BAR(DB_EMPTY + slot));
                        ++g;
                    }
                umma_commit_elect(DR_BAR(DB_XEMPTY + xs));
            }
            umma_commit_elect(DR_BAR(DB_D0FULL));
            DR_TL(lane == 0, 1);
            // ---- layer 1: A = layer-0 activations in TMEM, one complete K loop per 128-wide output half ----
#pragma unroll 1
            for (int h = 0; h < nh2; ++h) {
                const uint32_t d = h == 0 ? tD1a : tD1b;
#pragma unroll 1
                for (int c = 0; c < nc1; ++c) {
                    if (h == 0) {       // activations [64c, 64c + 64) are two epilogue batches
                        dr_wait(DR_BAR(DB_A1READY + 2 * c), ph);
                        dr_wait(DR_BAR(DB_A1READY + 2 * c + 1), ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        DR_TL(lane == 0 && c == 0, 2);
                    }
#pragma unroll 1
                    for (int we = 0; we < 2; ++we) {
                        if (we == 1 && sg == 0) continue;
                        const int slot = g % p.ns;
                        dr_wait(DR_BAR(DB_FULL + slot), (uint32_t)((g / p.ns) & 1));
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t bdesc = make_desc_sw128(ring0 + (uint32_t)slot * DR_TILE);
#pragma unroll
                        for (int j = 0; j < DR_KC / 16; ++j)
                            dr_umma_ts(d, tA1 + (uint32_t)(c * 32 + j * 8), bdesc + (uint64_t)(j * 2), we ? id_e128 : id_w128,
                                       (c | we | j) ? 1u : 0u);
                        umma_commit_elect(DR_BAR(DB_EMPTY + slot));
                        ++g;
                    }
                }
                umma_commit_elect(DR_BAR(DB_D1FULL + h));
                DR_TL(lane == 0, 3 + h);
            }
            // ---- head: A = layer-1 activations in TMEM ----
            {
                const uint32_t id_wh = dr_idesc(p.N3, 0), id_eh = dr_idesc(p.N3, sg < 0);
#pragma unroll 1
                for (int c = 0; c < nc2; ++c) {
                    dr_wait(DR_BAR(DB_A2READY + 2 * c), ph);
                    dr_wait(DR_BAR(DB_A2READY + 2 * c + 1), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    DR_TL(lane == 0 && c == 0, 5);
#pragma unroll 1
                    for (int we = 0; we < 2; ++we) {
                        if (we == 1 && sg == 0) continue;
                        const int slot = g % p.ns;
                        dr_wait(DR_BAR(DB_FULL + slot), (uint32_t)((g / p.ns) & 1));
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t bdesc = make_desc_sw128(ring0 + (uint32_t)slot * DR_TILE);
#pragma unroll
                        for (int j = 0; j < DR_KC / 16; ++j)
                            dr_umma_ts(tDH, tA2 + (uint32_t)(c * 32 + j * 8), bdesc + (uint64_t)(j * 2), we ? id_eh : id_wh,
                                       (c | we | j) ? 1u : 0u);
                        umma_commit_elect(DR_BAR(DB_EMPTY + slot));
                        ++g;
                    }
                }
                umma_commit_elect(DR_BAR(DB_DHFULL));
                DR_TL(lane == 0, 6);
            }
            __syncwarp();
            ph ^= 1u;
        }
    } else if (warp == 2) {
        // =============================== perturbed biases =======================================
        for (int u = 0; u < n_my; ++u) {
            const int work = (int)blockIdx.x + u * (int)gridDim.x;
            int m, tile;
            dr_item(p, work, m, tile);
            const int bb = u & 1;
            dr_wait(DR_BAR(DB_BEMPTY + bb), (uint32_t)((u >> 1) & 1) ^ 1u);
            const float sgf = p.sigma * (float)sign[m];
            const float* row = table_row_ptr(replicas, stride, idx[m]);
            float* bs = bias_s + bb * 768;
#pragma unroll
            for (int l = 0; l < 3; ++l)
                for (int i = lane; i < 256; i += 32) {
                    float v = 0.f;
                    if (i < p.nreal[l]) {
                        const int q = p.b_off[l] + i;
                        v = perturb1(theta[q], sgf, row[q]);
                    }
                    bs[l * 256 + i] = v;
                }
            __syncwarp();
            if (lane == 0) dr_arrive(DR_BAR(DB_BFULL + bb));
        }
    } else if (warp >= DR_CVT_WARP0) {
        // =============================== observation tile -> fp16 A operand =====================
        // 16 lanes cover the 256 bytes one observation row contributes to a chunk (coalesced), a warp instruction covers
        // two rows; the 8-byte piece of lane l16 lands at row * 128 + (((l16 >> 1) ^ (row & 7)) << 4) + (l16 & 1) * 8
        const int cw = warp - DR_CVT_WARP0, l16 = lane & 15, lr = lane >> 4;
        int xg = 0;
        for (int u = 0; u < n_my; ++u) {
            const int work = (int)blockIdx.x + u * (int)gridDim.x;
            int m, tile;
            dr_item(p, work, m, tile);
            const int e0i = tile * 128, ne = min(128, p.E - e0i);
            const float* ob = obs + ((int64_t)m * p.E + e0i) * p.K0;
#pragma unroll 1
            for (int c = 0; c < nc0; ++c, ++xg) {
                const int xs = xg % p.nx;
                const int k = c * DR_KC + l16 * 4;
                float4 v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int r = cw * 32 + i * 2 + lr;
                    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < ne && k < p.K0) v[i] = ldg_stream_f4(ob + (int64_t)r * p.K0 + k);      // K0 % 8 == 0: whole quads
                }
                dr_wait(DR_BAR(DB_XEMPTY + xs), (uint32_t)((xg / p.nx) & 1) ^ 1u);
                const uint32_t xd = xring0 + (uint32_t)xs * DR_TILE;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int r = cw * 32 + i * 2 + lr;
                    const uint32_t d = xd + (uint32_t)(r * 128 + (((l16 >> 1) ^ (r & 7)) << 4) + (l16 & 1) * 8);
                    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(d), "r"(dr_pack(v[i].x, v[i].y)), "r"(dr_pack(v[i].z, v[i].w)) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) dr_arrive(DR_BAR(DB_XFULL + xs));
                DR_TL(cw == 0 && lane == 0 && c == 0, 20);
                DR_TL(cw == 0 && lane == 0 && c == nc0 - 1, 21);
            }
        }
    } else if (warp >= DR_EPI_WARP0) {
        // =============================== epilogue warps =========================================
        const int ew = warp - DR_EPI_WARP0, q = ew & 3, set = ew >> 2;      // warp % 4 == q: its TMEM lane quarter
        const int gt = q * 32 + lane;                                      // observation row of this thread
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        uint32_t ph = 0;
        for (int u = 0; u < n_my; ++u) {
            const int work = (int)blockIdx.x + u * (int)gridDim.x;
            int m, tile;
            dr_item(p, work, m, tile);
            const int e0i = tile * 128, ne = min(128, p.E - e0i);
            const int bb = u & 1;
            const float* bs = bias_s + bb * 768;
            dr_wait(DR_BAR(DB_BFULL + bb), (uint32_t)((u >> 1) & 1));
            // ---- layer 0 accumulator -> layer-1 A operand ----
            DR_TL(ew == 0 && lane == 0, 19);
            dr_wait(DR_BAR(DB_D0FULL), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            DR_TL(ew == 0 && lane == 0, 8);
#pragma unroll 1
            for (int j = set; j < p.N1 / 32; j += 2) {
                uint32_t r[32], o[16];
                dr_tmem_ld32(tD0 + lane_sel + (uint32_t)(j * 32), r);
                const float4* b4 = reinterpret_cast<const float4*>(bs + j * 32);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 b = b4[i];
                    o[2 * i] = dr_pack(dr_tanh<APPROX>(__uint_as_float(r[4 * i]) + b.x), dr_tanh<APPROX>(__uint_as_float(r[4 * i + 1]) + b.y));
                    o[2 * i + 1] = dr_pack(dr_tanh<APPROX>(__uint_as_float(r[4 * i + 2]) + b.z), dr_tanh<APPROX>(__uint_as_float(r[4 * i + 3]) + b.w));
                }
                dr_tmem_st16(tA1 + lane_sel + (uint32_t)(j * 16), o);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) dr_arrive(DR_BAR(DB_A1READY + j));
            }
            DR_TL(ew == 0 && lane == 0, 9);
            // ---- layer 1 accumulator (two halves) -> layer-2 A operand ----
#pragma unroll 1
            for (int h = 0; h < nh2; ++h) {
                dr_wait(DR_BAR(DB_D1FULL + h), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                DR_TL(ew == 0 && lane == 0, 10 + 2 * h);
                const uint32_t d = h == 0 ? tD1a : tD1b;
#pragma unroll 1
                for (int jj = set; jj < 4; jj += 2) {
                    const int j = h * 4 + jj;
                    uint32_t r[32], o[16];
                    dr_tmem_ld32(d + lane_sel + (uint32_t)(jj * 32), r);
                    const float4* b4 = reinterpret_cast<const float4*>(bs + 256 + j * 32);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b = b4[i];
                        o[2 * i] = dr_pack(dr_tanh<APPROX>(__uint_as_float(r[4 * i]) + b.x), dr_tanh<APPROX>(__uint_as_float(r[4 * i + 1]) + b.y));
                        o[2 * i + 1] = dr_pack(dr_tanh<APPROX>(__uint_as_float(r[4 * i + 2]) + b.z), dr_tanh<APPROX>(__uint_as_float(r[4 * i + 3]) + b.w));
                    }
                    dr_tmem_st16(tA2 + lane_sel + (uint32_t)(j * 16), o);
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) dr_arrive(DR_BAR(DB_A2READY + j));
                }
                DR_TL(ew == 0 && lane == 0, 11 + 2 * h);
            }
            // ---- head (set 0 only): accumulator -> mean | std rows, staged and bulk-stored ----
            if (set == 0) {
                dr_wait(DR_BAR(DB_DHFULL), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                DR_TL(ew == 0 && lane == 0, 14);
                float* o = out + ((int64_t)m * p.E + e0i + gt) * p.nout;
                float* og = out + ((int64_t)m * p.E + e0i) * p.nout;
                const uint32_t obytes = (uint32_t)(ne * p.nout) * 4u;
                const bool bulk = (obytes & 15u) == 0 && (((uintptr_t)og) & 15) == 0;
                if (bulk) {
                    if (gt == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous block has left the staging rows
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                }
                float* os = ostage + gt * p.nout;
#pragma unroll 1
                for (int c = 0; c < p.N3; c += 16) {
                    float v[16];
                    tmem_ld16(tDH + lane_sel + (uint32_t)c, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float y = dr_tanh<APPROX>(v[i] + bs[512 + c + i]);
                        v[i] = c + i < p.A ? y : 0.55f + 0.45f * y;      // MapContinuousToAction (torch_helpers.py:20-25)
                    }
                    if (gt < ne) {
                        float* dst = bulk ? os : o;
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (c + i < p.nout) dst[c + i] = v[i];
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (bulk) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                    if (gt == 0) {
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(og), "r"(smem_u32(ostage)), "r"(obytes) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            }
            __syncwarp();
            DR_TL(ew == 0 && lane == 0, 15);
            if (lane == 0) dr_arrive(DR_BAR(DB_BEMPTY + bb));
            ph ^= 1u;
        }
        if (set == 0 && gt == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
#undef DR_BAR
#undef DR_TL
}

}  // namespace

static inline int64_t dr_stride16(int64_t size) { return (size + 64 + 63) / 64 * 64; }

// fp16 theta scratch: at least 16 384 halves (the resident theta image of the 64 x 64 kernel needs 9 216 whatever P is)
static inline int64_t dr_theta16_cap(int64_t n_params) { return n_params > 16384 ? n_params : 16384; }

extern "C" size_t dfd_table_scaled16_bytes(int64_t size, int64_t n_params) {
    return (size_t)8 * (size_t)dr_stride16(size) * 2 + dfd_align_up((size_t)dr_theta16_cap(n_params) * 2, 256) + 256;
}

int dfd_ws16_supported(const dfd_policy_desc* desc);

extern "C" int dfd_policy_direct_supported(const dfd_policy_desc* desc) {
    if (!desc || desc->kind != DFD_POLICY_MUJOCO) return 0;
    if (dfd_ws16_supported(desc)) return 1;
    const int K0 = desc->n_in, N1 = desc->h1, N2 = desc->h2, nout = 2 * desc->n_act;
    // whole 16-byte groups of halves along K in every layer (row pitch and layer offsets of the TMA maps), hidden widths in
    // whole 128-column accumulator halves, head within one MMA
    return (K0 % 8 == 0 && K0 >= 64 && (N1 == 128 || N1 == 256) && (N2 == 128 || N2 == 256) && nout <= 48) ? 1 : 0;
}

extern "C" int dfd_table_build_scaled16(dfd_ctx* ctx, const dfd_table* table, float sigma, int64_t n_params, void* buf,
                                        size_t bytes, dfd_stream stream) {
    DFD_CHECK_ARG(ctx && table && buf, "dfd_table_build_scaled16: NULL argument");
    DFD_CHECK_ARG(bytes >= dfd_table_scaled16_bytes(table->size, n_params), "dfd_table_build_scaled16: buffer too small");
    DFD_CHECK_ARG((((uintptr_t)buf) & 255) == 0, "dfd_table_build_scaled16: buffer must be 256-byte aligned");
    const int64_t s16 = dr_stride16(table->size);
    table_scaled16_kernel<<<ctx->sm_count * 8, 256, 0, (cudaStream_t)stream>>>(table->replicas, table->size, sigma, (__half*)buf, s16);
    DFD_LAUNCHED(ctx);
    ctx->scaled_src = table->replicas;
    ctx->scaled_sigma = sigma;
    ctx->scaled16 = buf;
    ctx->scaled16_stride = s16;
    ctx->theta16 = (char*)buf + (size_t)8 * (size_t)s16 * 2;
    ctx->theta16_cap = dr_theta16_cap(n_params);
    return 0;
}

extern "C" int dfd_table_drop_scaled16(dfd_ctx* ctx) {
    DFD_CHECK_ARG(ctx, "dfd_table_drop_scaled16: NULL context");
    ctx->scaled_src = nullptr;
    ctx->scaled16 = nullptr;
    ctx->theta16 = nullptr;
    ctx->theta16_cap = 0;
    return 0;
}

// returns -1 when this path does not serve the call (shape, or no scaled mirror of this table for this sigma)
int dfd_mlp_forward_direct_impl(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                                const int64_t* idx, const int8_t* sign, int n_members, float sigma, const float* obs,
                                int obs_per_member, float* out, int approx_tanh, cudaStream_t st) {
    if (getenv("DFD_TC_NO_DIRECT") || !dfd_policy_direct_supported(desc)) return -1;
    if (!ctx->scaled16 || ctx->scaled_src != table->replicas || ctx->scaled_sigma != sigma) return -1;
    if ((((uintptr_t)obs) & 15)) return -1;
    const int K0 = desc->n_in, N1 = desc->h1, N2 = desc->h2, nout = 2 * desc->n_act;
    DrParams p = {};
    p.K0 = K0; p.N1 = N1; p.N2 = N2; p.nout = nout; p.A = desc->n_act;
    p.N3 = (nout + 15) / 16 * 16;
    const int in_[3] = {K0, N1, N2}, outr[3] = {N1, N2, nout};
    int off = 0;
    for (int l = 0; l < 3; ++l) {
        p.w_off[l] = off; off += in_[l] * outr[l];
        p.b_off[l] = off; off += outr[l];
        p.kin[l] = in_[l];
        p.nreal[l] = outr[l];
        if (p.w_off[l] % 8) return -1;
    }
    p.P = off;
    if (p.P > ctx->theta16_cap) return -1;
    p.E = obs_per_member;
    p.tiles = (obs_per_member + 127) / 128;
    DFD_CHECK_ARG((int64_t)n_members * p.tiles < 2147483647LL, "direct MLP path: too many work items");
    p.n_work = n_members * p.tiles;
    p.sigma = sigma;
    p.pair_order = (n_members % 2 == 0 && !getenv("DFD_DR_NOPAIR")) ? 1 : 0;
    p.ns = getenv("DFD_DR_NS") ? atoi(getenv("DFD_DR_NS")) : DR_NS_MAX;
    p.nx = getenv("DFD_DR_NX") ? atoi(getenv("DFD_DR_NX")) : 2;
    if (p.ns < 2 || p.ns > DR_NS_MAX) p.ns = DR_NS_MAX;
    if (p.nx < 1 || p.nx > DR_NX_MAX) p.nx = 2;
    DrMaps maps;
    for (int l = 0; l < 3; ++l) {
        const int rows = l == 2 ? p.N3 : 128;
        DFD_CHECK_ARG(dr_map_w(&maps.w[l], ctx, p.w_off[l], in_[l], outr[l], rows) == 0,
                      "direct MLP path: cuTensorMapEncodeTiled failed for the layer-%d weights", l);
        DFD_CHECK_ARG(dr_map_e(&maps.e[l], ctx, in_[l], outr[l], rows) == 0,
                      "direct MLP path: cuTensorMapEncodeTiled failed for the layer-%d table rows", l);
    }
    if (dr_theta16(ctx, theta, p.P, st)) return 3;
    const size_t smem = (size_t)(p.ns + p.nx) * DR_TILE + (2 * 768 + 128 * (size_t)nout) * sizeof(float) + 1024;
    DFD_CHECK_ARG(smem <= 227 * 1024, "direct MLP path: %zu bytes of shared memory", smem);
    int grid = ctx->sm_count;
    if (grid > p.n_work) grid = p.n_work;
    long long* prof = nullptr;
    if (getenv("DFD_DR_PROF")) { cudaMalloc(&prof, (size_t)grid * 32 * 8); cudaMemset(prof, 0, (size_t)grid * 32 * 8); }
    if (approx_tanh) {
        DFD_CUDA(cudaFuncSetAttribute(mlp_forward_direct_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mlp_forward_direct_kernel<true><<<grid, DR_THREADS, smem, st>>>(p, maps, table->replicas, table->replica_stride, theta, idx, sign, obs, out, prof);
    } else {
        DFD_CUDA(cudaFuncSetAttribute(mlp_forward_direct_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mlp_forward_direct_kernel<false><<<grid, DR_THREADS, smem, st>>>(p, maps, table->replicas, table->replica_stride, theta, idx, sign, obs, out, prof);
    }
    DFD_LAUNCHED(ctx);
    if (prof) {
        cudaStreamSynchronize(st);
        static long long h[32];
        for (int cta = 7; cta < grid; cta += 60) {
            cudaMemcpy(h, prof + 32 * cta, sizeof(h), cudaMemcpyDeviceToHost);
            const long long t0 = h[0];
            fprintf(stderr, "[direct timeline] CTA %d unit 3 (cycles from L0 start) MMA: x0 ready %lld L0 issued %lld | a1 first %lld L1h0 issued %lld L1h1 issued %lld | a2 first %lld head issued %lld || "
                    "EPI: enter %lld d0full %lld epi0 done %lld | d1a %lld done %lld | d1b %lld done %lld | dh %lld end %lld || TMA: L0 first %lld L1 first %lld head first %lld || CVT: first %lld last %lld\n",
                    cta, h[7]-t0, h[1]-t0, h[2]-t0, h[3]-t0, h[4]-t0, h[5]-t0, h[6]-t0, h[19]-t0, h[8]-t0, h[9]-t0, h[10]-t0, h[11]-t0, h[12]-t0, h[13]-t0, h[14]-t0, h[15]-t0,
                    h[16]-t0, h[17]-t0, h[18]-t0, h[20]-t0, h[21]-t0);
        }
        cudaFree(prof);
    }
    return 0;
}
