// Atari perturbed forward on the tensor cores (policies/atari.py:35-51; perturbation worker/worker.py:28), precision >= 1.
//
// Every layer is linear in its weights, so a member's weights theta + s*sigma*eps are never built (csrc/direct_common.cuh):
// each contraction is   x.theta^T  (+/-)  x.(sigma*eps)^T   as tcgen05 kind::f16 MMAs whose WEIGHT operands arrive by TMA
// straight from the fp16 copy of theta and from the sigma-scaled fp16 mirror of the noise table.
//   conv 4->16 k8 s4   implicit GEMM: M = 400 output pixels (4 tiles of 128), N = 16, K = 256 in the weights' own order
//                      (c, ky, kx); the im2col tile is written by the CTA into the 128-byte swizzled K-major layout - one
//                      16-byte chunk = the 8 consecutive frame pixels (ky fixed) of a window row;
//   conv 16->32 k4 s2  M = 81 (one tile), N = 32, K = 256 (c, ky, kx); chunk = two 4-pixel window rows of the fp16 map;
//   Linear 2592->256   98 % of the parameters: "swap AB" - the WEIGHT tile [128 neurons x 64 k] is the A operand (TMA,
//                      7-slot ring of 16 KB tiles), the activations of up to 16 (member, observation) columns are the B
//                      operand (N = 16), accumulators [128 x 16] in TMEM: theta part and eps part separately, combined
//                      with the column's sign in the epilogue;
//   BN folds, Linear 256->A, softmax on the CUDA cores in fp32 with exactly perturbed parameters.
// One CTA evaluates the two members of an antithetic pair (members j and j + M/2): the eps tiles stream ONCE for both.
// fp16 operands (10-bit mantissa, the precision of tf32), fp32 accumulate; stated tolerance 2e-3 on the probabilities.
#include "direct_common.cuh"

namespace {

constexpr int AC_WORKERS = 512, AC_THREADS = AC_WORKERS + 32;
constexpr int FRAME = 4 * 84 * 84, A1N = 2592;
constexpr int AC_KB7 = 41;                  // 64-wide K boxes of the first Linear (2592 = 40.5 boxes; the tail is zero)
constexpr int AC_XT = AC_KB7 * 2048;        // x tiles: [16 columns x 64 k] per box
constexpr int AC_BUF = 65536;               // im2col tile [128 x 256] fp16
constexpr int AC_CW = 49152;                // conv weight tiles: W0 8 KB | E0 8 KB | W3 16 KB | E3 16 KB
constexpr int AC_NSLOT = (AC_BUF + AC_CW) / 16384;     // 7 ring slots for the first Linear
constexpr int AC_A0H = 16 * 400 * 2;
constexpr int AC_SCF = 96 + 4 * 256;        // floats: sc0 sh0 (16 + 16) sc3 sh3 (32 + 32) s8 sh8 [2][256]

// flat parameter offsets and BN buffer offsets (state_dict order)
constexpr int O_W0 = 0, O_B0 = 4096, O_G1 = 4112, O_BE1 = 4128, O_W3 = 4144, O_B3 = 12336, O_G4 = 12368,
              O_BE4 = 12400, O_W7 = 12432, O_B7 = 675984, O_G8 = 676240, O_BE8 = 676496, O_W10 = 676752;
constexpr int BU_M1 = 0, BU_V1 = 16, BU_M4 = 33, BU_V4 = 65, BU_M8 = 98, BU_V8 = 354;

struct AcMaps {
    CUtensorMap w0, e0, w3, e3, w7, e7;
};

enum { AB_CW = 0, AB_MMA = 1, AB_FULL = 2, AB_EMPTY = AB_FULL + AC_NSLOT, AB_COUNT = AB_EMPTY + AC_NSLOT };

__global__ void __launch_bounds__(AC_THREADS, 1)
atari_forward_tc_kernel(const __grid_constant__ AcMaps maps, const float* __restrict__ replicas, int64_t stride,
                        const float* __restrict__ theta, const float* __restrict__ bnbuf, const int64_t* __restrict__ idx,
                        const int8_t* __restrict__ sign, float sigma, const float* __restrict__ obs, int E, int A,
                        float* __restrict__ out, int n_members, int pair_mode) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[AB_COUNT];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const bool worker = tid < AC_WORKERS;
    const uint32_t bar0 = smem_u32(&bars[0]);
#define AC_BAR(i) (bar0 + 8u * (uint32_t)(i))
    const uint32_t sbase = smem_u32(smem_raw);
    const uint32_t xt0 = (sbase + 1023u) & ~1023u;
    uint8_t* xt = smem_raw + (xt0 - sbase);
    const uint32_t buf0 = xt0 + AC_XT;
    uint8_t* buf = xt + AC_XT;
    const uint32_t cw0 = buf0 + AC_BUF;
    __half* a0h = reinterpret_cast<__half*>(buf + AC_BUF + AC_CW);
    float* scf = reinterpret_cast<float*>(buf + AC_BUF + AC_CW + AC_A0H);
    float *sc0 = scf, *sh0 = scf + 16, *sc3 = scf + 32, *sh3 = scf + 64, *s8 = scf + 96, *sh8 = scf + 96 + 512;
    float* a2 = reinterpret_cast<float*>(buf);            // [16][256], aliases the ring once the first Linear is done
    float* lg = a2 + 16 * 256;                            // [16][32]

    const int nmem = pair_mode ? 2 : 1;
    const int m0 = (int)blockIdx.x;
    const int ms[2] = {m0, pair_mode ? m0 + (n_members >> 1) : m0};
    const int sgi[2] = {(int)sign[ms[0]], (int)sign[ms[1]]};
    const int64_t ids[2] = {idx[ms[0]], idx[ms[1]]};
    const float* rows[2] = {table_row_ptr(replicas, stride, ids[0]), table_row_ptr(replicas, stride, ids[1])};
    const bool shared_row = nmem == 2 && ids[0] == ids[1];
    const int nE = shared_row ? 1 : nmem;                 // distinct table rows whose eps tiles must stream

    // zero the x tiles (unused columns and the K tail must be zeros, not stale shared memory)
    for (int i = tid; i < AC_XT / 16; i += AC_THREADS) reinterpret_cast<uint4*>(xt)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(AC_BAR(AB_CW)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(AC_BAR(AB_MMA)));
        for (int s = 0; s < AC_NSLOT; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(AC_BAR(AB_FULL + s)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(AC_BAR(AB_EMPTY + s)));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t mph = 0, cwph = 0;             // parities of AB_MMA / AB_CW, advanced identically by every thread

    for (int mem = 0; mem < nmem; ++mem) {
        const int m = ms[mem];
        const float sg = sigma * (float)sgi[mem];
        const float* row = rows[mem];
        const bool load_e = mem == 0 || !shared_row;
        // ---- conv weight tiles by TMA: theta tiles once per CTA, eps tiles once per distinct table row ----
        if (warp == AC_WORKERS / 32 && lane == 0 && (mem == 0 || load_e)) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const uint32_t bytes = (mem == 0 ? 8192u + 16384u : 0u) + (load_e ? 8192u + 16384u : 0u);
            dr_expect_tx(AC_BAR(AB_CW), bytes);
            const int64_t s0 = ids[mem] + O_W0, s3 = ids[mem] + O_W3;
            for (int kb = 0; kb < 4; ++kb) {
                if (mem == 0) {
                    dr_tma_2d(cw0 + kb * 2048, &maps.w0, kb * 64, 0, AC_BAR(AB_CW));
                    dr_tma_2d(cw0 + 16384 + kb * 4096, &maps.w3, kb * 64, 0, AC_BAR(AB_CW));
                }
                if (load_e) {
                    dr_tma_4d(cw0 + 8192 + kb * 2048, &maps.e0, kb * 64, (int)(s0 >> 3), 0, (int)(s0 & 7), AC_BAR(AB_CW));
                    dr_tma_4d(cw0 + 32768 + kb * 4096, &maps.e3, kb * 64, (int)(s3 >> 3), 0, (int)(s3 & 7), AC_BAR(AB_CW));
                }
            }
        }
        // ---- BN folds of this member (conv bias and eval-mode BatchNorm -> per-channel scale / shift), exact fp32 ----
        if (tid < 16) {
            const float inv = 1.0f / sqrtf(bnbuf[BU_V1 + tid] + 1e-5f);
            const float s = perturb1(theta[O_G1 + tid], sg, row[O_G1 + tid]) * inv;
            sc0[tid] = s;
            sh0[tid] = (perturb1(theta[O_B0 + tid], sg, row[O_B0 + tid]) - bnbuf[BU_M1 + tid]) * s + perturb1(theta[O_BE1 + tid], sg, row[O_BE1 + tid]);
        } else if (tid >= 32 && tid < 64) {
            const int c = tid - 32;
            const float inv = 1.0f / sqrtf(bnbuf[BU_V4 + c] + 1e-5f);
            const float s = perturb1(theta[O_G4 + c], sg, row[O_G4 + c]) * inv;
            sc3[c] = s;
            sh3[c] = (perturb1(theta[O_B3 + c], sg, row[O_B3 + c]) - bnbuf[BU_M4 + c]) * s + perturb1(theta[O_BE4 + c], sg, row[O_BE4 + c]);
        } else if (tid >= 256 && tid < 512) {
            const int o = tid - 256;
            const float inv = 1.0f / sqrtf(bnbuf[BU_V8 + o] + 1e-5f);
            const float s = perturb1(theta[O_G8 + o], sg, row[O_G8 + o]) * inv;
            s8[mem * 256 + o] = s;
            sh8[mem * 256 + o] = (perturb1(theta[O_B7 + o], sg, row[O_B7 + o]) - bnbuf[BU_M8 + o]) * s + perturb1(theta[O_BE8 + o], sg, row[O_BE8 + o]);
        }
        if (mem == 0 || load_e) { dr_wait(AC_BAR(AB_CW), cwph); cwph ^= 1u; }
        __syncthreads();
        const uint32_t id16w = dr_idesc(16, 0), id16e = dr_idesc(16, sgi[mem] < 0);
        const uint32_t id32w = dr_idesc(32, 0), id32e = dr_idesc(32, sgi[mem] < 0);

        for (int e = 0; e < E; ++e) {
            const int n = mem * E + e;                    // column of this (member, observation)
            const float* fr = obs + ((int64_t)m * E + e) * FRAME;
            // ================= conv 4->16, k 8, s 4 =================
#pragma unroll 1
            for (int t = 0; t < 4; ++t) {
                // im2col tile t: task = (pixel row r, 16-byte chunk ch = c * 8 + ky) -> frame[c][4 oy + ky][4 ox .. 4 ox + 8)
                if (worker) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int task = i * AC_WORKERS + tid, r = task & 127, ch = task >> 7;
                        const int pix = t * 128 + r;
                        if (pix < 400) {
                            const int oy = pix / 20, ox = pix - oy * 20, c = ch >> 3, ky = ch & 7;
                            const float4* src = reinterpret_cast<const float4*>(fr + c * 7056 + (4 * oy + ky) * 84 + 4 * ox);
                            const float4 x0 = __ldg(src), x1 = __ldg(src + 1);
                            const uint32_t d = buf0 + (uint32_t)(c * 16384 + r * 128 + ((ky ^ (r & 7)) << 4));
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(dr_pack(x0.x, x0.y)), "r"(dr_pack(x0.z, x0.w)),
                                         "r"(dr_pack(x1.x, x1.y)), "r"(dr_pack(x1.z, x1.w)) : "memory");
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
                if (warp == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                    for (int we = 0; we < 2; ++we) {
                        if (we == 1 && sgi[mem] == 0) continue;
#pragma unroll 1
                        for (int kb = 0; kb < 4; ++kb) {
                            const uint64_t adesc = make_desc_sw128(buf0 + kb * 16384);
                            const uint64_t bdesc = make_desc_sw128(cw0 + (we ? 8192 : 0) + kb * 2048);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                dr_umma_ss(tmem + (uint32_t)(t * 16), adesc + (uint64_t)(j * 2), bdesc + (uint64_t)(j * 2), we ? id16e : id16w,
                                           (we | kb | j) ? 1u : 0u);
                        }
                    }
                    umma_commit_elect(AC_BAR(AB_MMA));
                }
                dr_wait(AC_BAR(AB_MMA), mph); mph ^= 1u;         // the tile has been read: the buffer may be rewritten
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (warp < 4) {                                   // accumulator -> scale / shift -> ReLU -> fp16 map [16][20][20]
                    float v[16];
                    tmem_ld16(tmem + lane_sel + (uint32_t)(t * 16), v);
                    const int pix = t * 128 + warp * 32 + lane;
                    if (pix < 400) {
#pragma unroll
                        for (int oc = 0; oc < 16; ++oc) a0h[oc * 400 + pix] = __float2half_rn(fmaxf(fmaf(v[oc], sc0[oc], sh0[oc]), 0.f));
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                }
            }
            __syncthreads();
            // ================= conv 16->32, k 4, s 2 =================
            // task = (pixel row r < 81, chunk ch): box b = ch >> 3 holds channels 4b .. 4b + 3; chunk j = ch & 7 = channel
            // 4b + (j >> 1), window rows ky = 2 (j & 1), 2 (j & 1) + 1, 4 pixels each
            if (worker) {
#pragma unroll 1
                for (int task = tid; task < 128 * 32; task += AC_WORKERS) {
                    const int r = task & 127, ch = task >> 7;
                    if (r < 81) {
                        const int oy = r / 9, ox = r - oy * 9, b = ch >> 3, j = ch & 7, c = 4 * b + (j >> 1), ky = 2 * (j & 1);
                        const uint32_t* s0 = reinterpret_cast<const uint32_t*>(a0h + c * 400 + (2 * oy + ky) * 20 + 2 * ox);
                        const uint32_t d = buf0 + (uint32_t)(b * 16384 + r * 128 + ((j ^ (r & 7)) << 4));
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(s0[0]), "r"(s0[1]), "r"(s0[10]), "r"(s0[11]) : "memory");
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (warp == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                for (int we = 0; we < 2; ++we) {
                    if (we == 1 && sgi[mem] == 0) continue;
#pragma unroll 1
                    for (int kb = 0; kb < 4; ++kb) {
                        const uint64_t adesc = make_desc_sw128(buf0 + kb * 16384);
                        const uint64_t bdesc = make_desc_sw128(cw0 + (we ? 32768 : 16384) + kb * 4096);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            dr_umma_ss(tmem + 64u, adesc + (uint64_t)(j * 2), bdesc + (uint64_t)(j * 2), we ? id32e : id32w, (we | kb | j) ? 1u : 0u);
                    }
                }
                umma_commit_elect(AC_BAR(AB_MMA));
            }
            dr_wait(AC_BAR(AB_MMA), mph); mph ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (warp < 4) {       // accumulator -> scale / shift -> ReLU -> column n of the x tiles, k = oc * 81 + pixel (flatten C,H,W)
                const int pix = warp * 32 + lane;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float v[16];
                    tmem_ld16(tmem + lane_sel + 64u + (uint32_t)(h * 16), v);
                    if (pix < 81) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int oc = h * 16 + i, k = oc * 81 + pix, kk = k & 63;
                            const __half hv = __float2half_rn(fmaxf(fmaf(v[i], sc3[oc], sh3[oc]), 0.f));
                            *reinterpret_cast<__half*>(xt + (k >> 6) * 2048 + n * 128 + (((kk >> 3) ^ (n & 7)) << 4) + (kk & 7) * 2) = hv;
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
            __syncthreads();
        }
    }
    // ================= Linear 2592 -> 256: weight tiles are the A operand, the x tiles the B operand =================
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the ring region was written through the generic proxy
    __syncthreads();
    const uint32_t tDW = tmem + 96u, tDE = tmem + 128u;
    if (warp == AC_WORKERS / 32) {
        if (lane == 0) {
            int g = 0;
#pragma unroll 1
            for (int kb = 0; kb < AC_KB7; ++kb)
#pragma unroll 1
                for (int part = 0; part < 1 + nE; ++part)
#pragma unroll 1
                    for (int mt = 0; mt < 2; ++mt, ++g) {
                        const int slot = g % AC_NSLOT;
                        dr_wait(AC_BAR(AB_EMPTY + slot), (uint32_t)((g / AC_NSLOT) & 1) ^ 1u);
                        dr_expect_tx(AC_BAR(AB_FULL + slot), 16384u);
                        const uint32_t dst = buf0 + (uint32_t)slot * 16384u;
                        if (part == 0) dr_tma_2d(dst, &maps.w7, kb * 64, mt * 128, AC_BAR(AB_FULL + slot));
                        else {
                            const int64_t s7 = ids[part - 1] + O_W7;
                            dr_tma_4d(dst, &maps.e7, kb * 64, (int)(s7 >> 3), mt * 128, (int)(s7 & 7), AC_BAR(AB_FULL + slot));
                        }
                    }
        }
    } else if (warp == 0) {
        const uint32_t id16 = dr_idesc(16, 0);
        int g = 0;
#pragma unroll 1
        for (int kb = 0; kb < AC_KB7; ++kb) {
            const uint64_t bdesc = make_desc_sw128(xt0 + (uint32_t)kb * 2048u);
#pragma unroll 1
            for (int part = 0; part < 1 + nE; ++part)
#pragma unroll 1
                for (int mt = 0; mt < 2; ++mt, ++g) {
                    const int slot = g % AC_NSLOT;
                    dr_wait(AC_BAR(AB_FULL + slot), (uint32_t)((g / AC_NSLOT) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t adesc = make_desc_sw128(buf0 + (uint32_t)slot * 16384u);
                    const uint32_t d = part == 0 ? tDW + (uint32_t)(mt * 16) : tDE + (uint32_t)(((part - 1) * 2 + mt) * 16);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        dr_umma_ss(d, adesc + (uint64_t)(j * 2), bdesc + (uint64_t)(j * 2), id16, (kb | j) ? 1u : 0u);
                    umma_commit_elect(AC_BAR(AB_EMPTY + slot));
                }
        }
        umma_commit_elect(AC_BAR(AB_MMA));
    }
    dr_wait(AC_BAR(AB_MMA), mph); mph ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // y = theta part + sign * eps part (+ bias, BN8, ReLU): thread = output neuron, 16 columns
    if (warp < 8) {
        const int mt = warp >> 2, o = mt * 128 + (warp & 3) * 32 + lane;
        float vw[16], ve0[16], ve1[16];
        tmem_ld16(tDW + lane_sel + (uint32_t)(mt * 16), vw);
        tmem_ld16(tDE + lane_sel + (uint32_t)(mt * 16), ve0);
        if (nE == 2) tmem_ld16(tDE + lane_sel + (uint32_t)((2 + mt) * 16), ve1);
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            const int mem = n >= E ? 1 : 0;
            if (n < nmem * E) {
                const float ev = (nE == 2 && mem == 1) ? ve1[n] : ve0[n];
                const float y = vw[n] + (float)sgi[mem] * ev;
                a2[n * 256 + o] = fmaxf(fmaf(y, s8[mem * 256 + o], sh8[mem * 256 + o]), 0.f);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    // ---- Linear 256 -> A (warp per action, exactly perturbed fp32 weights), softmax per column ----
    for (int a = warp; a < A && worker; a += AC_WORKERS / 32) {
        for (int mem = 0; mem < nmem; ++mem) {
            const float sg = sigma * (float)sgi[mem];
            const float* row = rows[mem];
            float w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = perturb1(theta[O_W10 + a * 256 + lane + 32 * j], sg, row[O_W10 + a * 256 + lane + 32 * j]);
            const float b = perturb1(theta[O_W10 + A * 256 + a], sg, row[O_W10 + A * 256 + a]);
            for (int e = 0; e < E; ++e) {
                const int n = mem * E + e;
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) s = fmaf(w[j], a2[n * 256 + lane + 32 * j], s);
                s = warp_sum(s);
                if (lane == 0) lg[n * 32 + a] = s + b;
            }
        }
    }
    __syncthreads();
    if (tid < nmem * E) {
        const int mem = tid >= E ? 1 : 0, e = tid - mem * E;
        const float* l = lg + tid * 32;
        float mx = -INFINITY;
        for (int a = 0; a < A; ++a) mx = fmaxf(mx, l[a]);
        float s = 0.f;
        for (int a = 0; a < A; ++a) s += expf(l[a] - mx);
        const float inv = 1.0f / s;
        float* o = out + ((int64_t)ms[mem] * E + e) * A;
        for (int a = 0; a < A; ++a) o[a] = expf(l[a] - mx) * inv;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
    }
#undef AC_BAR
}

}  // namespace

// returns -1 when this path does not serve the call (no scaled mirror of this table for this sigma, or too many
// observations per member for one pass)
int dfd_atari_forward_tc_impl(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                              const float* bn_buffers, const int64_t* idx, const int8_t* sign, int n_members, float sigma,
                              const float* obs, int obs_per_member, float* out, cudaStream_t st) {
    if (getenv("DFD_TC_NO_DIRECT")) return -1;
    if (!ctx->scaled16 || ctx->scaled_src != table->replicas || ctx->scaled_sigma != sigma) return -1;
    const int64_t P = dfd_policy_num_params(desc);
    if (P > ctx->theta16_cap) return -1;
    const int pair_mode = (n_members % 2 == 0 && obs_per_member <= 8) ? 1 : 0;
    if (!pair_mode && obs_per_member > 16) return -1;
    AcMaps maps;
    int rc = 0;
    rc |= dr_map_w(&maps.w0, ctx, O_W0, 256, 16, 16);
    rc |= dr_map_e(&maps.e0, ctx, 256, 16, 16);
    rc |= dr_map_w(&maps.w3, ctx, O_W3, 256, 32, 32);
    rc |= dr_map_e(&maps.e3, ctx, 256, 32, 32);
    rc |= dr_map_w(&maps.w7, ctx, O_W7, A1N, 256, 128);
    rc |= dr_map_e(&maps.e7, ctx, A1N, 256, 128);
    DFD_CHECK_ARG(rc == 0, "Atari tensor path: cuTensorMapEncodeTiled failed");
    if (dr_theta16(ctx, theta, P, st)) return 3;
    const size_t smem = (size_t)AC_XT + AC_BUF + AC_CW + AC_A0H + AC_SCF * sizeof(float) + 1024;
    DFD_CUDA(cudaFuncSetAttribute(atari_forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = pair_mode ? n_members / 2 : n_members;
    atari_forward_tc_kernel<<<grid, AC_THREADS, smem, st>>>(maps, table->replicas, table->replica_stride, theta, bn_buffers, idx, sign,
                                                            sigma, obs, obs_per_member, desc->n_act, out, n_members, pair_mode);
    DFD_LAUNCHED(ctx);
    return 0;
}
