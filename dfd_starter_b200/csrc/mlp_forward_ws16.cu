// Per-member MLP forward for the 64x64 MuJoCo nets (HalfCheetah-shaped 17-64-64-6 of BASELINE config 2; policies/mujoco.py:35-41,
// perturbation worker/worker.py:28) in the direct-from-table formulation (csrc/direct_common.cuh): a layer is linear in its
// weights, x.(theta + s*sigma*eps)^T = x.theta^T + s*(x.(sigma*eps)^T), so the perturbed weights are never built.
//   * theta tiles: an fp16 image of the three weight matrices (18 KB) is written ONCE per CTA (swizzled operand layout) and
//     stays resident for all its members;
//   * eps tiles: 18 KB per member straight from the sigma-scaled fp16 table mirror by TMA into a ring of member slots.  The
//     first layer's rows are K0 = 17 wide (34 bytes, not a legal TMA pitch): rows n = 8q + c form class c, 8 * K0 elements
//     apart, so eight boxes of 8 rows - one 1 KB swizzle atom each - make up the [64 x 64] tile in CLASS-MAJOR row order
//     (row c * 8 + q = neuron 8q + c); the theta image uses the same order and the epilogue undoes it for free (the
//     accumulator row is in registers: a column permutation is a renaming);
//   * the member's sign is the negate-A bit of the eps MMAs.
// No builder warps, no operand staging writes.  One persistent CTA per SM: warp 16 = TMA producer of the eps tiles (ten lanes
// issue one copy each), four MMA / epilogue groups of four warps, each with its own item in flight: observation row
// -> fp16 -> TMEM (A operand), tcgen05.mma kind::f16 (A from TMEM, B from shared memory) into a 64-column accumulator,
// epilogue accumulator -> + exactly perturbed fp32 bias -> tanh -> fp16x2 -> TMEM (the next layer's A operand), head ->
// MapContinuousToAction -> staged rows -> one bulk store.  TMEM per group: A0 16 | A 32 | D 64 columns (4 x 112 <= 512):
// fp16 A operands take half the columns of the tf32 kernel (mlp_forward_ws.cu), which is what lets a fourth item fly.
#include "direct_common.cuh"

namespace {

constexpr int W6_GROUPS = 4, W6_THREADS = W6_GROUPS * 128 + 32;
constexpr int W6_NE = 5;                       // eps ring: member slots of 18 KB
constexpr int W6_ESLOT = 8192 + 8192 + 2048;   // L0 tile (8 class atoms) | L1 tile | head tile
constexpr int W6_WBYTES = W6_ESLOT;            // resident theta image, same layout
constexpr int W6_TCOLS = 112;                  // TMEM columns per group: A0 [0,16) | A [16,48) | D [48,112)

struct W6Params {
    int K0, nout, A;
    int w_off1, w_off2, b_off0, b_off1, b_off2;
    int E, tiles, n_work, pair_order, obs_floats, ostage_floats;
    float sigma;
};
struct W6Maps {
    CUtensorMap e0, e1, e2;     // table mirror: class boxes {64, 1, 8, 1} | {64, 1, 64, 1} | {64, 1, 16, 1}
};

enum { WB_W = 0, WB_EFULL = 1, WB_EEMPTY = WB_EFULL + W6_NE, WB_OFULL = WB_EEMPTY + W6_NE, WB_OEMPTY = WB_OFULL + 2 * W6_GROUPS,
       WB_G = WB_OEMPTY + 2 * W6_GROUPS, WB_COUNT = WB_G + W6_GROUPS };

__device__ __forceinline__ void w6_item(const W6Params& p, int work, int& m, int& tile) {
    const int mm = p.tiles == 1 ? work : work / p.tiles;
    tile = work - mm * p.tiles;
    const int M = p.n_work / p.tiles;
    m = p.pair_order ? ((mm & 1) ? (M >> 1) + (mm >> 1) : (mm >> 1)) : mm;
}
__device__ __forceinline__ void w6_wait1(uint32_t bar, uint32_t parity) {      // one poller per warp
    if ((threadIdx.x & 31) == 0) dr_wait(bar, parity);
    __syncwarp();
}
__device__ __forceinline__ void w6_gsync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }
__device__ __forceinline__ void w6_tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}

// tanh of two pre-activations -> packed halves (`a` in the low half: the even k of the pair).  APPROX: round the pair to fp16
// first and take ONE tanh.approx.f16x2 (2^-10.99 relative, the class of the operand rounding that follows anyway): half the
// SFU operations of the fp32 form and no separate pack
template <bool APPROX>
__device__ __forceinline__ uint32_t w6_act2(float a, float b) {
    if (APPROX) {
        uint32_t h = dr_pack(a, b);
        asm("tanh.approx.f16x2 %0, %1;" : "=r"(h) : "r"(h));
        return h;
    } else {
        return dr_pack(tanh_fast(a), tanh_fast(b));
    }
}
__device__ __forceinline__ void w6_tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

template <bool APPROX>
__global__ void __launch_bounds__(W6_THREADS, 1)
mlp_forward_ws16_kernel(const W6Params p, const __grid_constant__ W6Maps maps, const float* __restrict__ replicas, int64_t stride,
                        const float* __restrict__ theta, const int64_t* __restrict__ idx, const int8_t* __restrict__ sign,
                        const float* __restrict__ obs, float* __restrict__ out, long long* __restrict__ prof) {
#define W6_TL(slot) do { if (prof && blockIdx.x == 7 && gt == 0 && g == 0 && u == 8) prof[slot] = clock64(); } while (0)
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[WB_COUNT];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int n_my = ((int)blockIdx.x < p.n_work) ? (p.n_work - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const uint32_t bar0 = smem_u32(&bars[0]);
#define W6_BAR(i) (bar0 + 8u * (uint32_t)(i))
    const uint32_t sraw = smem_u32(smem_raw);
    const uint32_t s0 = (sraw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (s0 - sraw);
    // [theta image 18 KB][eps ring NE x 18 KB][per group: 2 observation slots | staged output rows | 160 bias floats]
    const uint32_t w_s = s0, e_s = s0 + W6_WBYTES;
    float* gmem_f = reinterpret_cast<float*>(sm + W6_WBYTES + W6_NE * W6_ESLOT);
    const int gfloats = 2 * p.obs_floats + p.ostage_floats + 160;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(W6_BAR(WB_W)));
        for (int s = 0; s < W6_NE; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(W6_BAR(WB_EFULL + s)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(W6_BAR(WB_EEMPTY + s)));
        }
        for (int s = 0; s < 2 * W6_GROUPS; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(W6_BAR(WB_OFULL + s)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(W6_BAR(WB_OEMPTY + s)));
        }
        for (int g = 0; g < W6_GROUPS; ++g) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(W6_BAR(WB_G + g)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    dfd_grid_dependency_wait();        // programmatic dependent launch: the theta image / observations come from predecessors

    // resident theta image: fp16, the layout of an eps slot (W0 class-major rows, zero padded to 64 k | W1 | W2 zero padded to
    // 16 rows), written by the CTA itself in the 128-byte swizzle: 9 216 elements, once per kernel
    for (int i = tid; i < 9216 / 2; i += W6_THREADS) {
        const int e = 2 * i, row = e >> 6, k = e & 63;            // row of the [144 x 64] image, even k
        float v0 = 0.f, v1 = 0.f;
        if (row < 64) {
            const int n = 8 * (row & 7) + (row >> 3);
            if (k < p.K0) v0 = theta[n * p.K0 + k];
            if (k + 1 < p.K0) v1 = theta[n * p.K0 + k + 1];
        } else if (row < 128) {
            v0 = theta[p.w_off1 + (row - 64) * 64 + k]; v1 = theta[p.w_off1 + (row - 64) * 64 + k + 1];
        } else if (row - 128 < p.nout) {
            v0 = theta[p.w_off2 + (row - 128) * 64 + k]; v1 = theta[p.w_off2 + (row - 128) * 64 + k + 1];
        }
        *reinterpret_cast<uint32_t*>(sm + row * 128 + (((k >> 3) ^ (row & 7)) << 4) + (k & 7) * 2) = dr_pack(v0, v1);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if (warp == W6_GROUPS * 4) {
        // =============================== TMA producer ===========================================
        // lane 0 waits for the slots and arms the barriers, then eleven lanes issue one copy each (8 class boxes of layer
        // 0, layer 1, head, the observation tile): a single thread issuing them one after the other was the pace of the kernel
        for (int u = 0; u < n_my; ++u) {
            const int work = (int)blockIdx.x + u * (int)gridDim.x;
            int m, tile;
            w6_item(p, work, m, tile);
            const int es = u % W6_NE;
            if (lane == 0) {
                dr_wait(W6_BAR(WB_EEMPTY + es), (uint32_t)((u / W6_NE) & 1) ^ 1u);
                dr_expect_tx(W6_BAR(WB_EFULL + es), (uint32_t)W6_ESLOT);
            }
            __syncwarp();
            const uint32_t dst = e_s + (uint32_t)es * W6_ESLOT;
            const int64_t id = idx[m];
            if (lane < 8) {
                const int64_t s = id + (int64_t)p.K0 * lane;
                dr_tma_4d(dst + (uint32_t)lane * 1024u, &maps.e0, 0, (int)(s >> 3), 0, (int)(s & 7), W6_BAR(WB_EFULL + es));
            } else if (lane == 8) {
                const int64_t s1 = id + p.w_off1;
                dr_tma_4d(dst + 8192u, &maps.e1, 0, (int)(s1 >> 3), 0, (int)(s1 & 7), W6_BAR(WB_EFULL + es));
            } else if (lane == 9) {
                const int64_t s2 = id + p.w_off2;
                dr_tma_4d(dst + 16384u, &maps.e2, 0, (int)(s2 >> 3), 0, (int)(s2 & 7), W6_BAR(WB_EFULL + es));
            }
            __syncwarp();
        }
    } else {
        // =============================== MMA / epilogue groups ===================================
        const int g = warp >> 2, q = warp & 3, gt = q * 32 + lane;          // group, TMEM lane quarter, observation row
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const uint32_t tA0 = tmem + (uint32_t)(g * W6_TCOLS), tA = tA0 + 16u, tD = tA0 + 48u;
        float* gm = gmem_f + g * gfloats;
        float* ostage = gm + 2 * p.obs_floats;
        float* bias_s = ostage + p.ostage_floats;                           // [0,64) layer 0 | [64,128) layer 1 | [128,144) head
        uint32_t gph = 0;
        const uint32_t id64 = dr_idesc(64, 0), id16 = dr_idesc(16, 0);
        // the group fetches its own observation tiles, two items ahead: one bulk copy per item into the slot it has just read
        auto fetch_obs = [&](int uu) {
            if (uu < n_my) {
                int mm, tt;
                w6_item(p, (int)blockIdx.x + uu * (int)gridDim.x, mm, tt);
                const int e0 = tt * 128, nn = min(128, p.E - e0), o2 = (uu / W6_GROUPS) & 1;
                const uint32_t ob = (uint32_t)(nn * p.K0) * 4u;
                dr_expect_tx(W6_BAR(WB_OFULL + g * 2 + o2), ob);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(gm + o2 * p.obs_floats)), "l"(obs + ((int64_t)mm * p.E + e0) * p.K0), "r"(ob),
                             "r"(W6_BAR(WB_OFULL + g * 2 + o2)) : "memory");
            }
        };
        if (gt == 0) { fetch_obs(g); fetch_obs(g + W6_GROUPS); }
        // exactly perturbed biases (fp32, two roundings) of the group's NEXT item: two dependent global loads whose latency
        // hides under the current item
        float nbv = 0.f, nbh = 0.f;
        auto fetch_bias = [&](int uu) {
            if (uu < n_my) {
                int mm, tt;
                w6_item(p, (int)blockIdx.x + uu * (int)gridDim.x, mm, tt);
                const float sgf = p.sigma * (float)sign[mm];
                const float* trow = table_row_ptr(replicas, stride, idx[mm]);
                const int qo = gt < 64 ? p.b_off0 + gt : p.b_off1 + gt - 64;
                nbv = perturb1(theta[qo], sgf, trow[qo]);
                if (gt < p.nout) nbh = perturb1(theta[p.b_off2 + gt], sgf, trow[p.b_off2 + gt]);
            }
        };
        fetch_bias(g);
        for (int u = g; u < n_my; u += W6_GROUPS) {
            const int work = (int)blockIdx.x + u * (int)gridDim.x;
            int m, tile;
            w6_item(p, work, m, tile);
            const int e0i = tile * 128, ne = min(128, p.E - e0i);
            const int os = (u / W6_GROUPS) & 1, es = u % W6_NE;
            const int sg = (int)sign[m];
            const uint32_t id64e = dr_idesc(64, sg < 0), id16e = dr_idesc(16, sg < 0);
            W6_TL(0);
            const float bv = nbv, bh = nbh;                                     // fetched while the previous item computed
            // ---- observation row -> fp16 -> TMEM A0 (K0 <= 32 -> 16 columns; zero padded) ----
            w6_wait1(W6_BAR(WB_OFULL + g * 2 + os), (uint32_t)((u / (2 * W6_GROUPS)) & 1));
            W6_TL(1);
            {
                const float* orow = gm + os * p.obs_floats + gt * p.K0;
                uint32_t a0[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float x0 = 2 * j < p.K0 ? orow[2 * j] : 0.f, x1 = 2 * j + 1 < p.K0 ? orow[2 * j + 1] : 0.f;
                    a0[j] = dr_pack(x0, x1);
                }
                dr_tmem_st16(tA0 + lane_sel, a0);
            }
            bias_s[gt] = bv;
            if (gt < 16) bias_s[128 + gt] = bh;
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            w6_gsync(g);
            if (gt == 0) fetch_obs(u + 2 * W6_GROUPS);                          // the slot has been read by all four warps
            fetch_bias(u + W6_GROUPS);
            W6_TL(2);
            // ---- layer 0: D = A0 . W0^T (+/-) A0 . E0^T ----
            const uint32_t eb = e_s + (uint32_t)es * W6_ESLOT;
            if (q == 0) {
                w6_wait1(W6_BAR(WB_EFULL + es), (uint32_t)((u / W6_NE) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 2; ++j) dr_umma_ts(tD, tA0 + (uint32_t)(j * 8), make_desc_sw128(w_s) + (uint64_t)(j * 2), id64, j ? 1u : 0u);
                if (sg != 0) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) dr_umma_ts(tD, tA0 + (uint32_t)(j * 8), make_desc_sw128(eb) + (uint64_t)(j * 2), id64e, 1u);
                }
                umma_commit_elect(W6_BAR(WB_G + g));
            }
            w6_wait1(W6_BAR(WB_G + g), gph); gph ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            W6_TL(3);
            // ---- epilogues 0 / 1: accumulator -> + bias -> tanh -> fp16x2 -> A, 32 accumulator columns at a time ----
#pragma unroll 1
            for (int l = 0; l < 2; ++l) {
                const float* bs = bias_s + l * 64;
                uint32_t v[64];
                w6_tmem_ld32_nowait(tD + lane_sel, v);                    // both halves in flight, one wait
                w6_tmem_ld32_nowait(tD + lane_sel + 32u, v + 32);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (l == 0) {
                        // accumulator column c*8 + q holds neuron 8q + c (class-major rows of the layer-0 tiles): this half holds
                        // classes 4h .. 4h + 3, i.e. the neuron pairs (8q + 4h + 2t, + 1) = A columns 4q + 2h + t: undone for free
#pragma unroll
                        for (int qq = 0; qq < 8; ++qq) {
                            uint32_t o2[2];
#pragma unroll
                            for (int t = 0; t < 2; ++t) {
                                const int k0 = 8 * qq + 4 * h + 2 * t;
                                o2[t] = w6_act2<APPROX>(__uint_as_float(v[32 * h + (2 * t) * 8 + qq]) + bs[k0],
                                                        __uint_as_float(v[32 * h + (2 * t + 1) * 8 + qq]) + bs[k0 + 1]);
                            }
                            asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(tA + lane_sel + (uint32_t)(4 * qq + 2 * h)), "r"(o2[0]), "r"(o2[1]) : "memory");
                        }
                    } else {
                        uint32_t o[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            o[j] = w6_act2<APPROX>(__uint_as_float(v[32 * h + 2 * j]) + bs[32 * h + 2 * j], __uint_as_float(v[32 * h + 2 * j + 1]) + bs[32 * h + 2 * j + 1]);
                        asm volatile(
                            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(tA + lane_sel + (uint32_t)(16 * h)),
                            "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]), "r"(o[9]), "r"(o[10]),
                            "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]), "r"(o[15]) : "memory");
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                W6_TL(4 + 2 * l);
                w6_gsync(g);
                if (q == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (l == 0) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) dr_umma_ts(tD, tA + (uint32_t)(j * 8), make_desc_sw128(w_s + 8192u) + (uint64_t)(j * 2), id64, j ? 1u : 0u);
                        if (sg != 0) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) dr_umma_ts(tD, tA + (uint32_t)(j * 8), make_desc_sw128(eb + 8192u) + (uint64_t)(j * 2), id64e, 1u);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) dr_umma_ts(tD, tA + (uint32_t)(j * 8), make_desc_sw128(w_s + 16384u) + (uint64_t)(j * 2), id16, j ? 1u : 0u);
                        if (sg != 0) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) dr_umma_ts(tD, tA + (uint32_t)(j * 8), make_desc_sw128(eb + 16384u) + (uint64_t)(j * 2), id16e, 1u);
                        }
                        umma_commit_elect(W6_BAR(WB_EEMPTY + es));         // the member's eps tiles have been read
                    }
                    umma_commit_elect(W6_BAR(WB_G + g));
                }
                w6_wait1(W6_BAR(WB_G + g), gph); gph ^= 1u;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                W6_TL(5 + 2 * l);
            }
            // ---- head: mean | std rows (MapContinuousToAction, torch_helpers.py:20-25), staged and bulk-stored ----
            {
                float hv[16];
                tmem_ld16(tD + lane_sel, hv);
                float* og = out + ((int64_t)m * p.E + e0i) * p.nout;
                const uint32_t obytes = (uint32_t)(ne * p.nout) * 4u;
                const bool vec = (p.nout & 3) == 0 && (((uintptr_t)og) & 15) == 0;      // rows of whole float4s: contiguous in global memory
                const bool bulk = !vec && (obytes & 15u) == 0 && (((uintptr_t)og) & 15) == 0;
                if (vec) {
                    if (gt < ne) {
                        float y[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float t = dr_tanh<APPROX>(hv[i] + bias_s[128 + i]);
                            y[i] = i < p.A ? t : 0.55f + 0.45f * t;
                        }
                        float4* o4 = reinterpret_cast<float4*>(og + (int64_t)gt * p.nout);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (4 * i < p.nout) o4[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    w6_gsync(g);           // the bias buffer is rewritten by the next item
                    W6_TL(8);
                    continue;
                }
                if (bulk) {
                    if (gt == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous block has left the staging rows
                    w6_gsync(g);
                }
                float* dst = bulk ? ostage + gt * p.nout : og + (int64_t)gt * p.nout;
                if (gt < ne) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (i < p.nout) {
                            const float y = dr_tanh<APPROX>(hv[i] + bias_s[128 + i]);
                            dst[i] = i < p.A ? y : 0.55f + 0.45f * y;
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (bulk) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    w6_gsync(g);
                    if (gt == 0) {
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(og), "r"(smem_u32(ostage)), "r"(obytes) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                } else {
                    w6_gsync(g);           // the bias buffer is rewritten by the next item
                }
            }
            W6_TL(8);
        }
        if (gt == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
#undef W6_BAR
#undef W6_TL
}

// 4-D map over the table mirror for the 8-row class boxes of layer 0: {K0, starts, q (rows 8 * K0 elements apart), replica}
int w6_map_e_class(CUtensorMap* map, dfd_ctx* ctx, int K0) {
    dr_encode_fn encode = dr_encoder();
    if (!encode) return 1;
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const int64_t s16 = ctx->scaled16_stride;
    const cuuint64_t starts = (cuuint64_t)((s16 - (int64_t)64 * K0) / 8);
    const cuuint64_t dims[4] = {(cuuint64_t)K0, starts, 8, 8};
    const cuuint64_t strides[3] = {16, (cuuint64_t)K0 * 8 * 2, (cuuint64_t)s16 * 2};
    const cuuint32_t box[4] = {64, 1, 8, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, ctx->scaled16, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : 2;
}

}  // namespace

// 1 when the resident-theta direct kernel serves this MuJoCo shape (64 x 64 hidden, at most 32 inputs, 2A <= 16) AND it has
// been asked for (DFD_WS16=1): measured on B200 at C2 size it runs at the speed of mlp_forward_ws_kernel (24-28 us against
// 24.3 us: four items in flight, but a 13.6 k-cycle chain per item against 7.5 k - the four groups' tanh phases share one
// SFU and its TMEM reads are not pipelined under them), so the tf32 kernel with builder warps stays the default
int dfd_ws16_supported(const dfd_policy_desc* desc) {
    static const bool on = getenv("DFD_WS16") != nullptr;
    return (on && desc && desc->kind == DFD_POLICY_MUJOCO && desc->h1 == 64 && desc->h2 == 64 && desc->n_in >= 1 && desc->n_in <= 32 &&
            2 * desc->n_act <= 16) ? 1 : 0;
}

// returns -1 when this path does not serve the call (shape, alignment, or no scaled mirror of this table for this sigma)
int dfd_mlp_forward_ws16_impl(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                              const int64_t* idx, const int8_t* sign, int n_members, float sigma, const float* obs,
                              int obs_per_member, float* out, int approx_tanh, cudaStream_t st) {
    if (getenv("DFD_TC_NO_DIRECT") || !dfd_ws16_supported(desc)) return -1;
    if (!ctx->scaled16 || ctx->scaled_src != table->replicas || ctx->scaled_sigma != sigma) return -1;
    const int K0 = desc->n_in, nout = 2 * desc->n_act;
    // the observation tile of an item travels as one bulk copy: 16-byte aligned start and size
    if ((((uintptr_t)obs) & 15) || ((int64_t)obs_per_member * K0) % 4 || (128 * K0) % 4 || ((obs_per_member % 128) * K0) % 4) return -1;
    W6Params p = {};
    p.K0 = K0; p.nout = nout; p.A = desc->n_act;
    p.b_off0 = 64 * K0; p.w_off1 = p.b_off0 + 64; p.b_off1 = p.w_off1 + 4096; p.w_off2 = p.b_off1 + 64; p.b_off2 = p.w_off2 + nout * 64;
    if (p.w_off1 % 8 || p.w_off2 % 8) return -1;
    p.E = obs_per_member;
    p.tiles = (obs_per_member + 127) / 128;
    DFD_CHECK_ARG((int64_t)n_members * p.tiles < 2147483647LL, "resident-theta MLP path: too many work items");
    p.n_work = n_members * p.tiles;
    p.pair_order = (n_members % 2 == 0) ? 1 : 0;
    p.obs_floats = (128 * K0 + 3) / 4 * 4;
    p.ostage_floats = (128 * nout + 3) / 4 * 4;
    p.sigma = sigma;
    W6Maps maps;
    dr_encode_fn encode = dr_encoder();
    DFD_CHECK_ARG(encode != nullptr, "resident-theta MLP path: cuTensorMapEncodeTiled not available");
    int rc = 0;
    rc |= w6_map_e_class(&maps.e0, ctx, K0);
    rc |= dr_map_e(&maps.e1, ctx, 64, 64, 64);
    rc |= dr_map_e(&maps.e2, ctx, 64, nout, 16);
    DFD_CHECK_ARG(rc == 0, "resident-theta MLP path: cuTensorMapEncodeTiled failed");
    const size_t smem = (size_t)W6_WBYTES + (size_t)W6_NE * W6_ESLOT + (size_t)W6_GROUPS * (2 * p.obs_floats + p.ostage_floats + 160) * sizeof(float) + 1024;
    if (smem > 227 * 1024) return -1;
    int grid = ctx->sm_count;
    if (grid > p.n_work) grid = p.n_work;
    long long* prof = nullptr;
    if (getenv("DFD_W6_PROF")) { cudaMalloc(&prof, 16 * 8); cudaMemset(prof, 0, 16 * 8); }
    if (approx_tanh) {
        DFD_CUDA(cudaFuncSetAttribute(mlp_forward_ws16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DFD_CUDA(dfd_launch_pdl(mlp_forward_ws16_kernel<true>, dim3(grid), dim3(W6_THREADS), smem, st, p, maps, table->replicas, table->replica_stride,
                                theta, idx, sign, obs, out, prof));
    } else {
        DFD_CUDA(cudaFuncSetAttribute(mlp_forward_ws16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DFD_CUDA(dfd_launch_pdl(mlp_forward_ws16_kernel<false>, dim3(grid), dim3(W6_THREADS), smem, st, p, maps, table->replicas, table->replica_stride,
                                theta, idx, sign, obs, out, prof));
    }
    DFD_LAUNCHED(ctx);
    if (prof) {
        cudaStreamSynchronize(st);
        long long h[16];
        cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[ws16 timeline] CTA 7 group 0 item 8 (cycles): obs wait %lld | A0 + sync %lld | L0 mma %lld | epi0 %lld | sync + L1 mma %lld | epi1 %lld | sync + head mma %lld | head epilogue %lld | total %lld\n",
                h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5], h[7] - h[6], h[8] - h[7], h[8] - h[0]);
        cudaFree(prof);
    }
    return 0;
}
