// Warp-specialised STREAMING per-member MLP forward for the wide MuJoCo nets (hidden widths 64..256, e.g. the
// Humanoid-shaped 376-256-256-17 of BASELINE config 3; policies/mujoco.py:35-41, perturbation worker/worker.py:28).
// A member's weights (171 042 floats) do not fit on chip, so they stream through a shared-memory ring in K chunks:
//
//   builder warps (8)  two teams of four warps alternate chunks, so two chunks are always being built.  The theta tile
//                      [N x 32] of the chunk - and for layer 0 the observation chunk [128 x 32] (A) - arrive by TMA tensor
//                      copy (SWIZZLE_128B tensor maps) straight in the UMMA K-major operand layout, never touching L1;
//                      the builders load eps(member) from global memory (every row of the chunk is one 128-byte line: a
//                      quarter-warp reads one full line, all 16 loads of a thread in flight before the first use) and add
//                      s*sigma*eps IN PLACE with the reference's two roundings, tf32;
//   MMA warp (1 lane)  tcgen05.mma kind::tf32, M = 128 observations, N = layer width, 2 instructions per chunk;
//                      layer 0 takes A from shared memory, layers 1 and 2 take A from TENSOR MEMORY;
//   epilogue warps (4) TMEM accumulator -> + perturbed bias -> tanh -> tf32 -> the same TMEM columns (the next
//                      layer's A operand), 32 columns at a time, each batch published on its own mbarrier so the next
//                      layer's MMAs start on the first 32 activations while the rest are still being computed.
// TMEM: region R1 (columns 0-255) = layer-0 accumulator / layer-1 A operand / head accumulator,
//       region R2 (columns 256-511) = layer-1 accumulator / layer-2 A operand.
// The builders run ahead of the MMA warp by the depth of the ring, across layer and member boundaries.
#include "tc_common.cuh"
#include <cuda.h>

namespace {

constexpr int ST_KC = 32;                 // K columns per chunk: one 128-byte swizzle atom row
constexpr int ST_TEAMS = 2;               // builder teams: team t builds chunks t, t+2, ... so two chunks' loads are in flight
constexpr int ST_TEAM_WARPS = 4;
constexpr int ST_NS = 4;                  // ring stages (max; the launcher may use fewer to leave L1 for in-flight loads)
constexpr int ST_NBUILD = 256;
constexpr int ST_NEPI = 128;               // 4 warps, one per TMEM lane quarter (and per SM sub-partition: the SFU is per sub-partition)
constexpr int ST_THREADS = ST_NBUILD + ST_NEPI + 32;
constexpr int ST_WCHUNK = 256 * ST_KC;    // floats
constexpr int ST_ACHUNK = 128 * ST_KC;
constexpr int ST_STAGE = ST_WCHUNK + ST_ACHUNK;

struct StParams {
    int K0, N1, N2, nout, N3, A;
    int w_off[3], b_off[3], kin[3], nreal[3], npad[3], nchunk[3];
    int E, tiles, n_work, pair_order, prefetch, ns, head_fused;
    int64_t P;
    float sigma;
};

// theta tiles (one map per layer: [n_out rows, k_in columns] row-major, box 32 columns x npad rows) and observation chunks
// ([members * E rows, K0 columns], box 32 x 128), both SWIZZLE_128B: the copy lands in the operand layout directly and
// never passes through L1.  Out-of-bounds rows / columns (k >= k_in, padded rows) are filled with zeros.
struct StMaps {
    CUtensorMap w[3];
    CUtensorMap obs;
};

__device__ __forceinline__ void st_tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void st_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// work item -> (member, tile).  pair_order: consecutive work items are the two members of an antithetic pair
// ([plus | minus] batches: members j and j + M/2 share their table row), so the CTAs b and b+1 stream the same eps
// row at the same time and HBM serves it once.
__device__ __forceinline__ void st_item(const StParams& p, int work, int& m, int& tile) {
    const int mm = p.tiles == 1 ? work : work / p.tiles;
    tile = work - mm * p.tiles;
    const int M = p.n_work / p.tiles;
    m = p.pair_order ? ((mm & 1) ? (M >> 1) + (mm >> 1) : (mm >> 1)) : mm;
}

__device__ __forceinline__ float st_tf32(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }

template <bool APPROX>
__device__ __forceinline__ float st_tanh(float x) {
    if (APPROX) {
        float y;
        asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    } else {
        return tanh_fast(x);
    }
}

__device__ __forceinline__ void st_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(200000u)
            : "memory");
        if (!ok && ++spins > (1u << 22)) __trap();
    } while (!ok);
}
__device__ __forceinline__ void st_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void st_umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void st_tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// barrier slots
enum { SB_FULL = 0, SB_EMPTY = 4, SB_DFULL = 8, SB_HREADY = 11 /* [2 layers][8 chunks] */, SB_R1FREE = 27, SB_BFULL = 28, SB_BEMPTY = 30, SB_TFULL = 32, SB_COUNT = 36 };

template <bool APPROX>
__global__ void __launch_bounds__(ST_THREADS, 1)
mlp_forward_stream_kernel(const StParams p, const __grid_constant__ StMaps maps, const float* __restrict__ replicas, int64_t stride, const float* __restrict__ theta,
                          const int64_t* __restrict__ idx, const int8_t* __restrict__ sign, const float* __restrict__ obs,
                          float* __restrict__ out, long long* __restrict__ prof) {
#define ST_TL(cond, slot) do { if (prof && (cond) && u == 3) prof[(size_t)blockIdx.x * 32 + (slot)] = clock64(); } while (0)
    extern __shared__ __align__(128) float smem[];   // [NS stages: W chunk | A chunk][2 x 3 x 256 bias]
    __shared__ __align__(8) uint64_t bars[SB_COUNT];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int n_my = ((int)blockIdx.x < p.n_work) ? (p.n_work - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const uint32_t bar0 = smem_u32(&bars[0]);
    // stages are SWIZZLE_128B operands: 1024-byte aligned atoms (the launcher allocates 1 KB of slack)
    const uint32_t smem_raw = smem_u32(smem);
    const uint32_t smem0 = (smem_raw + 1023u) & ~1023u;
    float* bias_s = smem + ((smem0 - smem_raw) >> 2) + p.ns * ST_STAGE;
    float* ostage = bias_s + 2 * 768;            // [128 x nout] output rows of the member being finished (16-byte aligned)
#define ST_BAR(i) (bar0 + 8u * (uint32_t)(i))

    if (tid == 0) {
        for (int s = 0; s < ST_NS; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ST_BAR(SB_FULL + s)), "r"(ST_TEAM_WARPS));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ST_BAR(SB_EMPTY + s)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ST_BAR(SB_TFULL + s)));
        }
        for (int l = 0; l < 3; ++l) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ST_BAR(SB_DFULL + l)));
        for (int j = 0; j < 16; ++j) asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(ST_BAR(SB_HREADY + j)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ST_BAR(SB_R1FREE)), "r"(ST_NEPI / 32));
        for (int b = 0; b < 2; ++b) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ST_BAR(SB_BFULL + b)), "r"(ST_NBUILD / 32));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ST_BAR(SB_BEMPTY + b)), "r"(ST_NEPI / 32));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t tR1 = tmem, tR2 = tmem + 256u;

    if (warp < ST_NBUILD / 32) {
        // =============================== builders ===============================================
        // lane -> (row-in-item li = lane >> 3, 16-byte chunk lc = lane & 7): a warp item is 4 rows x 128 contiguous bytes
        // (4 full lines per request); row r's chunk lc lands at r * 128 + ((lc ^ (r & 7)) << 4) - the 128-byte swizzle
        const int li = lane >> 3, lc = lane & 7;
        const int team = warp / ST_TEAM_WARPS, tw = warp % ST_TEAM_WARPS;
        int g = 0;           // running chunk number over layers and members: stage = g % ns, team = g % ST_TEAMS
        for (int u = 0; u < n_my; ++u) {
            const int work = (int)blockIdx.x + u * (int)gridDim.x;
            int m, tile;
            st_item(p, work, m, tile);
            const int e0i = tile * 128, ne = min(128, p.E - e0i);
            const float sg = p.sigma * (float)sign[m];
            const float* row = table_row_ptr(replicas, stride, idx[m]);
            if (p.prefetch && tid == 0 && u + 1 < n_my) {      // next member's eps row and observation tile -> L2
                int nm, nt;
                st_item(p, work + (int)gridDim.x, nm, nt);
                l2_prefetch(table_row_ptr(replicas, stride, idx[nm]), (size_t)p.P * 4);
                l2_prefetch(obs + ((int64_t)nm * p.E + nt * 128) * p.K0, (size_t)min(128, p.E - nt * 128) * p.K0 * 4);
            }
            // perturbed biases of the three layers into bias buffer (u & 1)
            {
                const int bb = u & 1;
                st_wait(ST_BAR(SB_BEMPTY + bb), (uint32_t)((u >> 1) & 1) ^ 1u);
                float* bs = bias_s + bb * 768;
#pragma unroll
                for (int l = 0; l < 3; ++l) {
                    float v = 0.f;
                    if (tid < p.nreal[l]) {
                        const int q = p.b_off[l] + tid;
                        v = perturb1(theta[q], sg, row[q]);
                    }
                    bs[l * 256 + tid] = v;
                }
                __syncwarp();
                if (lane == 0) st_arrive(ST_BAR(SB_BFULL + bb));
            }
#pragma unroll
            for (int l = 0; l < 3; ++l) {      // unrolled: every p.xxx[l] is a compile-time constant-bank read
                const int kin = p.kin[l], nreal = p.nreal[l], npad = p.npad[l], nchunk = p.nchunk[l];
                const float* ep_l = row + p.w_off[l];
                if (l == 2 && p.head_fused) {
                    // the whole head layer ([npad x kin] <= 48 KB) as ONE ring entry: kin / 32 swizzle atoms of npad rows,
                    // one tensor copy each, instead of kin / 32 dependent chunk round trips of a few KB
                    if (g % ST_TEAMS == team) {
                        const int s = g % p.ns;
                        st_wait(ST_BAR(SB_EMPTY + s), (uint32_t)((g / p.ns) & 1) ^ 1u);
                        const uint32_t Wd = smem0 + 4u * (uint32_t)(s * ST_STAGE);
                        const int natoms = kin / ST_KC, ri = npad >> 2, items = natoms * ri;
                        const uint32_t atom_bytes = (uint32_t)npad * 128u;
                        if (tw == 0 && lane == 0) {
                            st_expect_tx(ST_BAR(SB_TFULL + s), (uint32_t)natoms * atom_bytes);
                            for (int a = 0; a < natoms; ++a)
                                st_tma_2d(Wd + (uint32_t)a * atom_bytes, &maps.w[2], a * ST_KC, 0, ST_BAR(SB_TFULL + s));
                        }
                        int atom0 = 0, ritem0 = tw;          // item t = h * 4 + tw -> (atom, row item) = (t / ri, t % ri); ri >= 4
#pragma unroll 1
                        for (int t0 = 0; t0 < items; t0 += 16 * ST_TEAM_WARPS) {
                            float4 e[16];
                            int atom = atom0, ritem = ritem0;
#pragma unroll
                            for (int h = 0; h < 16; ++h) {
                                const int n = ritem * 4 + li, k = atom * ST_KC + lc * 4;
                                e[h] = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (atom < natoms && n < nreal) e[h] = ldg_stream_f4(ep_l + n * kin + k);
                                ritem += ST_TEAM_WARPS;
                                if (ritem >= ri) { ritem -= ri; ++atom; }
                            }
                            if (t0 == 0) st_wait(ST_BAR(SB_TFULL + s), (uint32_t)((g / p.ns) & 1));
                            atom = atom0; ritem = ritem0;
#pragma unroll
                            for (int h = 0; h < 16; ++h) {
                                const int n = ritem * 4 + li;
                                if (atom < natoms) {
                                    const uint32_t d = Wd + (uint32_t)atom * atom_bytes + (uint32_t)(n * 128 + ((lc ^ (n & 7)) << 4));
                                    float4 a;
                                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(d));
                                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(d),
                                                 "f"(st_tf32(perturb1(a.x, sg, e[h].x))), "f"(st_tf32(perturb1(a.y, sg, e[h].y))),
                                                 "f"(st_tf32(perturb1(a.z, sg, e[h].z))), "f"(st_tf32(perturb1(a.w, sg, e[h].w)))
                                                 : "memory");
                                }
                                ritem += ST_TEAM_WARPS;
                                if (ritem >= ri) { ritem -= ri; ++atom; }
                            }
                            atom0 = atom; ritem0 = ritem;
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) st_arrive(ST_BAR(SB_FULL + s));
                    }
                    ++g;
                    continue;
                }
#pragma unroll 1
                for (int c = 0; c < nchunk; ++c, ++g) {
                    if (g % ST_TEAMS != team) continue;
                    const int s = g % p.ns;
                    st_wait(ST_BAR(SB_EMPTY + s), (uint32_t)((g / p.ns) & 1) ^ 1u);
                    const uint32_t Wd = smem0 + 4u * (uint32_t)(s * ST_STAGE);
                    const uint32_t Ad = Wd + 4u * ST_WCHUNK;
                    const int k = c * ST_KC + lc * 4;
                    // theta tile (and, for layer 0, the observation chunk) by tensor copy straight into the stage; the
                    // builders then add s*sigma*eps IN PLACE: eps from global memory (the only loads left on the L1 path),
                    // theta from the swizzled position they write back to
                    if (tw == 0 && lane == 0) {
                        const uint32_t bytes = (uint32_t)npad * 128u + (l == 0 ? 128u * 128u : 0u);
                        st_expect_tx(ST_BAR(SB_TFULL + s), bytes);
                        st_tma_2d(Wd, &maps.w[l], c * ST_KC, 0, ST_BAR(SB_TFULL + s));
                        if (l == 0) st_tma_2d(Ad, &maps.obs, c * ST_KC, m * p.E + e0i, ST_BAR(SB_TFULL + s));
                    }
                    // ALL of the chunk's eps loads (16 x 16 bytes per thread for N = 256) are in flight before the first use:
                    // one memory round trip per chunk
                    constexpr int NI = 16;
                    float4 e[NI];
#pragma unroll
                    for (int h = 0; h < NI; ++h) {
                        const int n = (h * ST_TEAM_WARPS + tw) * 4 + li;
                        e[h] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (n < nreal && k < kin) e[h] = ldg_stream_f4(ep_l + n * kin + k);      // kin % 4 == 0: whole quads
                    }
                    st_wait(ST_BAR(SB_TFULL + s), (uint32_t)((g / p.ns) & 1));    // theta tile (and observations) landed
#pragma unroll
                    for (int h = 0; h < NI; ++h) {
                        const int n = (h * ST_TEAM_WARPS + tw) * 4 + li;
                        if (n < npad) {
                            const uint32_t d = Wd + (uint32_t)(n * 128 + ((lc ^ (n & 7)) << 4));
                            float4 a;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(d));
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(d),
                                         "f"(st_tf32(perturb1(a.x, sg, e[h].x))), "f"(st_tf32(perturb1(a.y, sg, e[h].y))),
                                         "f"(st_tf32(perturb1(a.z, sg, e[h].z))), "f"(st_tf32(perturb1(a.w, sg, e[h].w)))
                                         : "memory");
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) st_arrive(ST_BAR(SB_FULL + s));
                }
            }
        }
    } else if (warp < (ST_NBUILD + ST_NEPI) / 32) {
        // =============================== epilogue warps =========================================
        const int q = (warp - ST_NBUILD / 32) & 3;
        const int gt = q * 32 + lane;                 // observation row of this thread
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        uint32_t pd[3] = {0, 0, 0};
        for (int u = 0; u < n_my; ++u) {
            const int work = (int)blockIdx.x + u * (int)gridDim.x;
            int m, tile;
            st_item(p, work, m, tile);
            const int e0i = tile * 128, ne = min(128, p.E - e0i);
            const int bb = u & 1;
            const float* bs = bias_s + bb * 768;
            st_wait(ST_BAR(SB_BFULL + bb), (uint32_t)((u >> 1) & 1));
#pragma unroll
            for (int l = 0; l < 2; ++l) {
                const uint32_t treg = l == 0 ? tR1 : tR2;
                const int N = p.npad[l];
                ST_TL(gt == 0, 8 + 2 * l);
                st_wait(ST_BAR(SB_DFULL + l), pd[l]); pd[l] ^= 1u;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                ST_TL(gt == 0, 9 + 2 * l);
#pragma unroll 1
                for (int c0 = 0; c0 < N; c0 += 32) {
                    uint32_t r[32];
                    st_tmem_ld32(treg + lane_sel + (uint32_t)c0, r);
                    const float4* b4 = reinterpret_cast<const float4*>(bs + l * 256 + c0);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b = b4[i];
                        r[4 * i + 0] = __float_as_uint(st_tf32(st_tanh<APPROX>(__uint_as_float(r[4 * i + 0]) + b.x)));
                        r[4 * i + 1] = __float_as_uint(st_tf32(st_tanh<APPROX>(__uint_as_float(r[4 * i + 1]) + b.y)));
                        r[4 * i + 2] = __float_as_uint(st_tf32(st_tanh<APPROX>(__uint_as_float(r[4 * i + 2]) + b.z)));
                        r[4 * i + 3] = __float_as_uint(st_tf32(st_tanh<APPROX>(__uint_as_float(r[4 * i + 3]) + b.w)));
                    }
                    st_tmem_st32(treg + lane_sel + (uint32_t)c0, r);
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) st_arrive(ST_BAR(SB_HREADY + l * 8 + (c0 >> 5)));   // these 32 activations are an A chunk now
                }
            }
            // head: accumulator in R1 columns [0, N3).  The member's [ne x nout] output block is contiguous in global
            // memory, so the rows are staged in shared memory and leave with ONE bulk store: per-thread row stores touch 32
            // lines per instruction (34 of them per thread: ~9 000 cycles of the SM's single L1 wavefront queue per member,
            // taken from the builders' loads, and the next member's layer 0 waited that long for region R1).  R1 is
            // released as soon as the accumulator is in registers.
            ST_TL(gt == 0, 12);
            st_wait(ST_BAR(SB_DFULL + 2), pd[2]); pd[2] ^= 1u;
            ST_TL(gt == 0, 13);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                float* o = out + ((int64_t)m * p.E + e0i + gt) * p.nout;
                float* og = out + ((int64_t)m * p.E + e0i) * p.nout;
                const uint32_t obytes = (uint32_t)(ne * p.nout) * 4u;
                const bool bulk = (obytes & 15u) == 0 && (((uintptr_t)og) & 15) == 0;
                if (bulk) {
                    // the previous member's bulk store must have finished READING the staging rows
                    if (gt == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                }
                float* os = ostage + gt * p.nout;
#pragma unroll 1
                for (int c = 0; c < p.N3; c += 16) {
                    float v[16];
                    tmem_ld16(tR1 + lane_sel + (uint32_t)c, v);
                    if (c + 16 >= p.N3) {       // last read of R1: the next member's layer 0 may overwrite it
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) st_arrive(ST_BAR(SB_R1FREE));
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float y = st_tanh<APPROX>(v[i] + bs[512 + c + i]);
                        v[i] = c + i < p.A ? y : 0.55f + 0.45f * y;   // MapContinuousToAction
                    }
                    if (gt < ne) {
                        float* dst = bulk ? os : o;
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (c + i < p.nout) dst[c + i] = v[i];
                    }
                }
                if (bulk) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                    if (gt == 0) {
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(og),
                                     "r"(smem_u32(ostage)), "r"(obytes)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            }
            __syncwarp();
            ST_TL(gt == 0, 14);
            if (lane == 0) st_arrive(ST_BAR(SB_BEMPTY + bb));
        }
        if (gt == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // the last block has left shared memory
    } else {
        // =============================== MMA issue warp =========================================
        int g = 0;
        uint32_t ph = 0, pr1 = 0;
        for (int u = 0; u < n_my; ++u) {
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                const uint32_t idesc = make_idesc_tf32(p.npad[l]);
                const uint32_t d_tmem = l == 1 ? tR2 : tR1;
                const uint32_t a_tmem = l == 1 ? tR1 : tR2;
                ST_TL(lane == 0, 2 * l);
                if (l == 0 && u > 0) { st_wait(ST_BAR(SB_R1FREE), pr1); pr1 ^= 1u; }   // head of the previous member read out
                ST_TL(lane == 0 && l == 0, 6);
                if (l == 2 && p.head_fused) {
                    const int s = g % p.ns, natoms = p.kin[2] / ST_KC;
                    st_wait(ST_BAR(SB_FULL + s), (uint32_t)((g / p.ns) & 1));
                    const uint32_t Wd = smem0 + 4u * (uint32_t)(s * ST_STAGE);
#pragma unroll 1
                    for (int a = 0; a < natoms; ++a) {
                        st_wait(ST_BAR(SB_HREADY + 8 + a), ph);          // activations [32a, 32a + 32) of layer 1 are in TMEM
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t bdesc = make_desc_sw128(Wd + (uint32_t)a * (uint32_t)p.npad[2] * 128u);
#pragma unroll
                        for (int j = 0; j < ST_KC / 8; ++j)
                            st_umma_ts(d_tmem, a_tmem + (uint32_t)(a * ST_KC + j * 8), bdesc + (uint64_t)(j * 2), idesc, (a | j) ? 1u : 0u);
                    }
                    umma_commit_elect(ST_BAR(SB_EMPTY + s));
                    umma_commit_elect(ST_BAR(SB_DFULL + 2));
                    __syncwarp();
                    ++g;
                    ST_TL(lane == 0, 2 * l + 1);
                    continue;
                }
#pragma unroll 1
                for (int c = 0; c < p.nchunk[l]; ++c, ++g) {
                    const int s = g % p.ns;
                    if (l > 0 && (c * ST_KC) % 32 == 0)
                        st_wait(ST_BAR(SB_HREADY + (l - 1) * 8 + (c * ST_KC) / 32), ph);   // these activations are in TMEM
                    st_wait(ST_BAR(SB_FULL + s), (uint32_t)((g / p.ns) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t Wd = smem0 + 4u * (uint32_t)(s * ST_STAGE);
                    // SWIZZLE_128B K-major operands: 8-row atoms of 1024 bytes; a K step of 8 tf32 (32 bytes) inside the atom
                    // advances the start-address field by 2
                    const uint64_t bdesc = make_desc_sw128(Wd);
                    if (l == 0) {
                        const uint64_t adesc = make_desc_sw128(Wd + 4u * ST_WCHUNK);
#pragma unroll
                        for (int j = 0; j < ST_KC / 8; ++j)
                            umma_tf32_elect(d_tmem, adesc + (uint64_t)(j * 2), bdesc + (uint64_t)(j * 2), idesc, (c | j) ? 1u : 0u);
                    } else {
#pragma unroll
                        for (int j = 0; j < ST_KC / 8; ++j)
                            st_umma_ts(d_tmem, a_tmem + (uint32_t)(c * ST_KC + j * 8), bdesc + (uint64_t)(j * 2), idesc, (c | j) ? 1u : 0u);
                    }
                    umma_commit_elect(ST_BAR(SB_EMPTY + s));
                    if (c == p.nchunk[l] - 1) umma_commit_elect(ST_BAR(SB_DFULL + l));
                    __syncwarp();
                }
                ST_TL(lane == 0, 2 * l + 1);
            }
            ph ^= 1u;          // every activation barrier completes once per member
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
#undef ST_BAR
#undef ST_TL
}

}  // namespace

typedef CUresult (*st_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// fp32 [rows, cols] row-major -> boxes of 32 columns x box_rows rows in the 128-byte swizzle; 0 on success
static int st_make_map(CUtensorMap* map, const float* base, uint64_t cols, uint64_t rows, uint32_t box_rows) {
    static st_encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
            cudaGetLastError();
            return 1;
        }
        encode = (st_encode_fn)fn;
    }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)ST_KC, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : 2;
}

// returns -1 when the shape is not served by this kernel (the caller falls back to the generic tcgen05 kernel)
int dfd_mlp_forward_stream_impl(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                                const int64_t* idx, const int8_t* sign, int n_members, float sigma, const float* obs,
                                int obs_per_member, float* out, int approx_tanh, cudaStream_t st) {
    const int K0 = desc->n_in, N1 = desc->h1, N2 = desc->h2, nout = 2 * desc->n_act;
    if (getenv("DFD_TC_NO_STREAM")) return -1;
    // whole 16-byte quads along K in every layer, hidden widths in whole 64-column halves, head within one MMA
    if (K0 % 4 || N1 % 64 || N2 % 64 || N1 > 256 || N2 > 256 || N1 < 64 || N2 < 64 || nout > 64) return -1;
    if ((((uintptr_t)theta) & 15) || (((uintptr_t)obs) & 15)) return -1;
    StParams p = {};
    p.K0 = K0; p.N1 = N1; p.N2 = N2; p.nout = nout; p.A = desc->n_act;
    p.N3 = (nout + 15) / 16 * 16;
    const int in_[3] = {K0, N1, N2}, outr[3] = {N1, N2, nout}, outp[3] = {N1, N2, p.N3};
    int off = 0;
    for (int l = 0; l < 3; ++l) {
        p.w_off[l] = off; off += in_[l] * outr[l];
        p.b_off[l] = off; off += outr[l];
        p.kin[l] = in_[l];
        p.nreal[l] = outr[l];
        p.npad[l] = outp[l];
        p.nchunk[l] = (in_[l] + ST_KC - 1) / ST_KC;
        if (p.w_off[l] % 4) return -1;
    }
    p.P = off;
    p.E = obs_per_member;
    p.tiles = (obs_per_member + 127) / 128;
    DFD_CHECK_ARG((int64_t)n_members * p.tiles < 2147483647LL, "tcgen05 MLP path: too many work items");
    p.n_work = n_members * p.tiles;
    p.sigma = sigma;
    p.pair_order = (n_members % 2 == 0 && !getenv("DFD_ST_NOPAIR")) ? 1 : 0;
    p.prefetch = getenv("DFD_ST_NOPF") ? 0 : 1;
    // the head layer as one ring entry when it fits a stage (whole 32-column atoms, atoms 1024-byte aligned)
    p.head_fused = (!getenv("DFD_ST_NOFUSE") && N2 % ST_KC == 0 && p.N3 % 8 == 0 && p.N3 >= 16 && (size_t)p.N3 * N2 * 4 <= (size_t)ST_STAGE * 4) ? 1 : 0;
    // ns x 48 KB + biases - measured on B200:
    // the bytes of global loads in flight (and with them the builders' throughput) scale with the L1 that is left
    p.ns = getenv("DFD_ST_NS") ? atoi(getenv("DFD_ST_NS")) : 3;
    if (p.ns < 3 || p.ns > ST_NS) p.ns = 3;
    const size_t smem = ((size_t)p.ns * ST_STAGE + 2 * 768 + 128 * (size_t)nout) * sizeof(float) + 1024;   // + alignment slack of the swizzled stages
    StMaps maps;
    for (int l = 0; l < 3; ++l)
        DFD_CHECK_ARG(st_make_map(&maps.w[l], theta + p.w_off[l], (uint64_t)p.kin[l], (uint64_t)p.nreal[l], (uint32_t)p.npad[l]) == 0,
                      "tcgen05 MLP path: cuTensorMapEncodeTiled failed for the layer-%d weights", l);
    DFD_CHECK_ARG(st_make_map(&maps.obs, obs, (uint64_t)K0, (uint64_t)n_members * (uint64_t)obs_per_member, 128u) == 0,
                  "tcgen05 MLP path: cuTensorMapEncodeTiled failed for the observations");
    int grid = ctx->sm_count;
    if (grid > p.n_work) grid = p.n_work;
    long long* prof = nullptr;
    if (getenv("DFD_ST_PROF")) { cudaMalloc(&prof, (size_t)grid * 32 * 8); cudaMemset(prof, 0, (size_t)grid * 32 * 8); }
    if (approx_tanh) {
        DFD_CUDA(cudaFuncSetAttribute(mlp_forward_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mlp_forward_stream_kernel<true><<<grid, ST_THREADS, smem, st>>>(p, maps, table->replicas, table->replica_stride, theta, idx, sign, obs, out, prof);
    } else {
        DFD_CUDA(cudaFuncSetAttribute(mlp_forward_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mlp_forward_stream_kernel<false><<<grid, ST_THREADS, smem, st>>>(p, maps, table->replicas, table->replica_stride, theta, idx, sign, obs, out, prof);
    }
    DFD_LAUNCHED(ctx);
    if (prof) {
        cudaStreamSynchronize(st);
        static long long h[32];
        cudaMemcpy(h, prof + 32 * 7, sizeof(h), cudaMemcpyDeviceToHost);
        const long long t0 = h[0];
        fprintf(stderr, "[stream timeline] CTA 7 unit 3 (cycles from L0 start): MMA: L0 start %lld (r1free seen %lld) issued %lld | L1 start %lld issued %lld | L2 start %lld issued %lld || EPI: wait0 %lld got %lld | wait1 %lld got %lld | wait2 %lld got %lld | end %lld\n", h[0]-t0, h[6]-t0, h[1]-t0, h[2]-t0, h[3]-t0, h[4]-t0, h[5]-t0, h[8]-t0, h[9]-t0, h[10]-t0, h[11]-t0, h[12]-t0, h[13]-t0, h[14]-t0);
        cudaFree(prof);
    }
    return 0;
}
