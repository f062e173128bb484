// Warp-specialised, software-pipelined per-member MLP forward for the 64x64 MuJoCo nets
// (policies/mujoco.py:35-41 + utils/torch_helpers.py:20-25; perturbation worker/worker.py:28).
//
// One persistent CTA per SM, 8 builder warps + up to three MMA/epilogue groups of 4 warps (640 threads x 96 registers):
//   * producer duty (no warp of its own): cp.async.bulk (TMA bulk copy) of eps(member) - one contiguous,
//     16-byte aligned slice of a table replica - and of the member's observation tile into a shared-memory
//     ring, completion on an mbarrier (expect_tx).  The LAST builder warp to finish reading a ring slot
//     (shared-memory counter) refills it with the item NE ahead, so the copies run NE items ahead;
//   * builder warps (8): own theta in REGISTERS for the whole kernel (each thread always builds the same
//     elements), read eps from the ring, form theta + s*sigma*eps (one FMA: see tf32_fma), round to
//     tf32 and write the UMMA K-major canonical B operands of all three layers into one of NST operand
//     stages (bank-conflict-free diagonal lane mapping);
//   * two or three epilogue/MMA groups (4 warps each; three whenever 3 x 168 TMEM columns suffice).  Every A operand lives in TENSOR MEMORY: thread r
//     owns row r of the 128-observation tile, copies its observation row from the group's own TMA-fed ring
//     into TMEM (tcgen05.st), one elected lane issues tcgen05.mma (kind::tf32, A from TMEM, B from shared
//     memory), then each thread reads its accumulator row, applies tanh, rounds to tf32 and writes it back
//     IN PLACE - the next layer's A operand.  Activations never touch shared memory.  While one group waits
//     for its MMA the other group's tanh work fills the SFU pipe.
// Biases ride inside the GEMM: every A operand carries a constant-one column and the matching B column is
// the (perturbed) bias, so the epilogue is tanh + store only.
// TANH_APPROX = true uses the single-instruction tanh.approx.f32 (2^-11 relative, the same class as the
// tf32 operand rounding); false uses ex2/rcp (~1e-6 absolute) at twice the SFU cost.
#include "tc_common.cuh"

namespace {

constexpr int WS_HID = 64;        // hidden width served by this kernel
constexpr int WS_KH = 72;         // K of layers 1 and 2: 64 activations + the bias/ones column block
constexpr int WS_MAXG = 3;         // MMA/epilogue groups (runtime: 2 or 3)
constexpr int WS_NBUILD = 256;
constexpr int WS_THREADS = WS_NBUILD + 128 * WS_MAXG;
// TMEM columns per group: A0 (a0c) | A1/D1 72 | A2/D2 72 | D3.  Two groups: a0c = 40, D3 = 32 own columns (2 x 216);
// three groups: a0c = 24 and D3 (<= 16 columns) aliases the head of A1/D1, dead once layer 1 has run (3 x 168 = 504 <= 512).

struct WsParams {
    int K0, K0p, nkq, nout, N3, A;
    int w_off1, w_off2, b_off0, b_off1, b_off2;
    int E, tiles, n_work, nst, ne, obs_vec, ng, a0c, tcols, d3_off;
    int st_floats, o_w0, o_w1, o_w2;
    int o_ring, ring_floats, eps_floats, o_obs, obs_floats;
    float sigma;
};

// tf32 operands: tcgen05.mma kind::tf32 reads sign, exponent and the upper 10 mantissa bits of each 32-bit
// operand and ignores the low 13 bits, so adding half a tf32 ulp (one integer add) turns that truncation into
// round-to-nearest, ties away - the same value cvt.rna.tf32.f32 produces, without its conversion-pipe cost.
__device__ __forceinline__ float tf32_rn(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }

// operand element of the tensor path: theta + s*sigma*eps rounded to tf32.  One FMA instead of the reference's mul-then-add:
// the two differ by at most one fp32 ulp, 2^13 times below the tf32 rounding applied right after (the exact fp32
// path and dfd_perturb_members keep the bit-exact two-rounding form of worker/worker.py:28).
__device__ __forceinline__ float tf32_fma(float theta, float sg, float eps) { return tf32_rn(fmaf(sg, eps, theta)); }

template <bool APPROX>
__device__ __forceinline__ float ws_tanh(float x) {
    if (APPROX) {
        float y;
        asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    } else {
        return tanh_fast(x);
    }
}

// mbarrier wait that parks the warp in hardware (suspend-time hint) instead of spinning through the issue
// slots the other roles need; a lost arrival still faults instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_park(uint32_t bar, uint32_t parity) {
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

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// A operand from tensor memory (128 lanes x K columns of 32-bit elements), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__device__ __forceinline__ void umma_tf32_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
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

__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                 "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// barrier slots
enum { B_WFULL = 0, B_WEMPTY = 3, B_MMA = 6, B_EFULL = 9, B_OFULL = 13, B_COUNT = 19 };

__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, float x) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory"); }

template <bool APPROX>
__global__ void __launch_bounds__(WS_THREADS, 1)
mlp_forward_ws_kernel(const WsParams p, const float* __restrict__ replicas, int64_t stride, const float* __restrict__ theta,
                      const int64_t* __restrict__ idx, const int8_t* __restrict__ sign, const float* __restrict__ obs,
                      float* __restrict__ out, long long* __restrict__ prof) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bars[B_COUNT];
    __shared__ uint32_t tmem_base_s;
    __shared__ float sg_s[4];
    __shared__ int ring_cnt[4];

    const int tid = threadIdx.x, lane = tid & 31;
    // the warp index through a shuffle: the compiler then knows it is warp-uniform, keeps the role branches
    // and everything derived from them (stage bases, descriptors) on the uniform datapath
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int n_my = ((int)blockIdx.x < p.n_work) ? (p.n_work - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const uint32_t bar0 = smem_u32(&bars[0]);
    const uint32_t smem0 = smem_u32(smem);
#define WS_BAR(i) (bar0 + 8u * (uint32_t)(i))
// debugging aid (DFD_WS_PROF=1): clock64 stamps of one steady-state item per role
#define WS_STAMP(cond, slot) do { } while (0)
// (compiled in with -DDFD_WS_TIMELINE only: the group chain is instruction-latency bound, ~5 cycles per instruction,
// and eleven guarded stamps per item are not free)
#ifdef DFD_WS_TIMELINE
#define WS_TL(cond, item, slot) do { if (prof && (cond) && (item) >= 0 && (item) < 16) prof[(size_t)blockIdx.x * 256 + (item) * 16 + (slot)] = clock64(); } while (0)
#else
#define WS_TL(cond, item, slot) do { } while (0)
#endif

    if (tid < 4) ring_cnt[tid] = 0;
    if (tid == 0) {
        for (int s = 0; s < 3; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(WS_BAR(B_WFULL + s)), "r"(WS_NBUILD / 32));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(WS_BAR(B_WEMPTY + s)));
        }
        for (int s = 0; s < 4; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(WS_BAR(B_EFULL + s)));
        for (int s = 0; s < 2 * WS_MAXG; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(WS_BAR(B_OFULL + s)));
        for (int s = 0; s < WS_MAXG; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(WS_BAR(B_MMA + s)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // first NE eps rows: requested before anything else so the HBM latency overlaps the prologue
        for (int kk = 0; kk < p.ne && kk < n_my; ++kk) {
            const int work = (int)blockIdx.x + kk * (int)gridDim.x;
            const int m = p.tiles == 1 ? work : work / p.tiles;
            sg_s[kk] = p.sigma * (float)sign[m];
            const uint32_t eb = (uint32_t)p.eps_floats * 4u;
            mbar_expect_tx(WS_BAR(B_EFULL + kk), eb);
            bulk_g2s(smem0 + 4u * (uint32_t)(p.o_ring + kk * p.ring_floats), table_row_ptr(replicas, stride, idx[m]), eb,
                     WS_BAR(B_EFULL + kk));
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // zero the operand stages once (padding rows / columns stay zero for the whole kernel) and the zero word
    // at the tail of every eps ring slot (the source of every "no such element" in the builders)
    for (int i = tid; i < p.nst * p.st_floats; i += (int)blockDim.x) smem[i] = 0.f;
    if (tid < p.ne) smem[p.o_ring + tid * p.ring_floats + p.eps_floats] = 0.f;
    // launched with programmatic stream serialisation: everything above (barriers, the first eps copies - they read
    // only the table and the index / sign arrays -, TMEM allocation, zeroing) overlaps the predecessor's tail; theta
    // and the observations are read below
    dfd_grid_dependency_wait();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    if (warp < WS_NBUILD / 32) {   // warp order: builders 0-7, groups 8-11, 12-15 (, 16-19)
        // =============================== builders ===============================================
        const int bt = tid;
        const int li = lane & 7, lq = lane >> 3;
        // --- fixed element ownership: theta into registers, source / destination byte offsets precomputed.
        // W1 / W2: 8x8 blocks of (row, k-quad) walked along diagonals so that both the 16-byte ring reads
        // (bank = k-quad) and the 16-byte canonical stores (bank = row & 7) are conflict-free.
        float4 th1[4], th2[2], th0[2];
        uint32_t s1[4], d1[4], s2[2], d2[2], d0[2], src0[2][4];
        bool v2[2], v0[2];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int blk = warp * 2 + (b >> 1), rg = blk >> 1, h = blk & 1, c = (b & 1) * 4 + lq;
            const int r = rg * 8 + li, kq = h * 8 + ((li + c) & 7);
            const int so = p.w_off1 + r * WS_HID + kq * 4;
            s1[b] = 4u * (uint32_t)so;
            d1[b] = 4u * (uint32_t)(p.o_w1 + rg * (WS_KH * 8) + kq * 32 + li * 4);
            th1[b] = *reinterpret_cast<const float4*>(theta + so);
        }
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int rg = warp >> 1, h = warp & 1, c = b * 4 + lq;
            const int r = rg * 8 + li, kq = h * 8 + ((li + c) & 7);
            const int so = p.w_off2 + r * WS_HID + kq * 4;
            v2[b] = r < p.nout;
            s2[b] = v2[b] ? 4u * (uint32_t)so : 0u;
            d2[b] = 4u * (uint32_t)(p.o_w2 + rg * (WS_KH * 8) + kq * 32 + li * 4);
            th2[b] = v2[b] ? *reinterpret_cast<const float4*>(theta + so) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // W0 items: (row < 64, k-quad < nkq), item j -> r8 = j & 7 fastest so quarter-warps store 128 contiguous
        // bytes.  Element k < K0: weight; k == K0: the bias (rides in the GEMM against the ones column of A);
        // k > K0: theta = 0 and the source is the ring slot's zero word, so the loop below has no special cases.
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int j = bt + WS_NBUILD * b, q8 = j >> 3, rg = q8 / p.nkq, kq = q8 - rg * p.nkq, r = rg * 8 + (j & 7);
            v0[b] = rg < 8;
            d0[b] = 4u * (uint32_t)(p.o_w0 + rg * (p.K0p * 8) + kq * 32 + (j & 7) * 4);
            float t[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int kk = kq * 4 + c;
                int so = p.eps_floats;
                t[c] = 0.f;
                if (v0[b] && kk <= p.K0) {
                    so = kk < p.K0 ? r * p.K0 + kk : p.b_off0 + r;
                    t[c] = theta[so];
                }
                src0[b][c] = 4u * (uint32_t)so;
            }
            th0[b] = make_float4(t[0], t[1], t[2], t[3]);
        }
        // bias columns of W1 (threads 0-63) and W2 (threads 64..64+nout-1): element (r, k = 64)
        int bsrc = -1;
        uint32_t bdst = 0;
        if (bt < 64) { bsrc = p.b_off1 + bt; bdst = 4u * (uint32_t)(p.o_w1 + (bt >> 3) * (WS_KH * 8) + 16 * 32 + (bt & 7) * 4); }
        else if (bt < 64 + p.nout) { bsrc = p.b_off2 + bt - 64; bdst = 4u * (uint32_t)(p.o_w2 + ((bt - 64) >> 3) * (WS_KH * 8) + 16 * 32 + ((bt - 64) & 7) * 4); }
        const float thb = bsrc >= 0 ? theta[bsrc] : 0.f;
        const uint32_t bsrc_b = 4u * (uint32_t)(bsrc >= 0 ? bsrc : p.eps_floats);

        // producer duty: (row pointer, signed sigma) of item k + NE are fetched by lane 0 of every builder warp
        // at the top of iteration k, so whichever warp turns out to be the last reader of the slot can refill
        // it without waiting on global memory
        auto produce = [&](int slot_, const float* row, float sgv) {
            sg_s[slot_] = sgv;
            const uint32_t eb = (uint32_t)p.eps_floats * 4u;
            mbar_expect_tx(WS_BAR(B_EFULL + slot_), eb);
            bulk_g2s(smem0 + 4u * (uint32_t)(p.o_ring + slot_ * p.ring_floats), row, eb, WS_BAR(B_EFULL + slot_));
        };
        int s = 0, slot = 0;
        uint32_t pw = 0, pe = 0;
        // raw (idx, sign) of item k + NE + 1 are loaded at the top of iteration k and only converted at the top
        // of iteration k + 1, so the dependent arithmetic never waits on global memory
        long long idx_raw = 0;
        int sign_raw = 0;
        auto fetch_raw = [&](int kk) {
            if (lane == 0 && kk < n_my) {
                const int work = (int)blockIdx.x + kk * (int)gridDim.x;
                const int m = p.tiles == 1 ? work : work / p.tiles;
                idx_raw = idx[m];
                sign_raw = sign[m];
            }
        };
        fetch_raw(p.ne);
        for (int k = 0; k < n_my; ++k) {
            const float* row_n = table_row_ptr(replicas, stride, idx_raw);
            const float sg_n = p.sigma * (float)sign_raw;
            fetch_raw(k + p.ne + 1);
            WS_TL(bt == 32, k, 8);
            mbar_wait_park(WS_BAR(B_EFULL + slot), pe);           // eps landed
            WS_TL(bt == 32, k, 9);
            mbar_wait_park(WS_BAR(B_WEMPTY + s), pw ^ 1u);        // stage drained by item k - nst
            WS_TL(bt == 32, k, 10);
            const uint32_t S = smem0 + 4u * (uint32_t)(s * p.st_floats);
            const uint32_t eps = smem0 + 4u * (uint32_t)(p.o_ring + slot * p.ring_floats);
            const float sg = sg_s[slot];
            // ---- phase A: every read of the ring first, phase B: perturb, round, store
            // (two rounds, so at most 17 ring values are live at a time: 96-register budget)
            {
                float4 e1[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) e1[b] = lds128(eps + s1[b]);
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    sts128(S + d1[b], tf32_fma(th1[b].x, sg, e1[b].x), tf32_fma(th1[b].y, sg, e1[b].y),
                           tf32_fma(th1[b].z, sg, e1[b].z), tf32_fma(th1[b].w, sg, e1[b].w));
            }
            float4 e2[2];
            float ew[2][4];
#pragma unroll
            for (int b = 0; b < 2; ++b) e2[b] = lds128(eps + s2[b]);
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int c = 0; c < 4; ++c) ew[b][c] = lds32(eps + src0[b][c]);
            const float ebias = lds32(eps + bsrc_b);
#pragma unroll
            for (int b = 0; b < 2; ++b)
                if (v2[b])
                    sts128(S + d2[b], tf32_fma(th2[b].x, sg, e2[b].x), tf32_fma(th2[b].y, sg, e2[b].y),
                           tf32_fma(th2[b].z, sg, e2[b].z), tf32_fma(th2[b].w, sg, e2[b].w));
#pragma unroll
            for (int b = 0; b < 2; ++b)
                if (v0[b])
                    sts128(S + d0[b], tf32_fma(th0[b].x, sg, ew[b][0]), tf32_fma(th0[b].y, sg, ew[b][1]),
                           tf32_fma(th0[b].z, sg, ew[b][2]), tf32_fma(th0[b].w, sg, ew[b][3]));
            if (bsrc >= 0) sts32(S + bdst, tf32_fma(thb, sg, ebias));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(WS_BAR(B_WFULL + s));
                if (atomicAdd(&ring_cnt[slot], 1) == WS_NBUILD / 32 - 1) {   // last reader of the slot refills it
                    ring_cnt[slot] = 0;
                    if (k + p.ne < n_my) produce(slot, row_n, sg_n);
                }
            }
            if (++s == p.nst) { s = 0; pw ^= 1u; }
            if (++slot == p.ne) { slot = 0; pe ^= 1u; }
            WS_TL(bt == 32, k, 11);
        }
    } else {
        // =============================== MMA + epilogue groups ==================================
        const int g = (warp - WS_NBUILD / 32) >> 2, q = warp & 3;
        const int gt = tid & 127;                 // row of this thread inside the tile
        const uint32_t mbar = WS_BAR(B_MMA + g);
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const uint32_t tA0 = tmem + (uint32_t)(g * p.tcols), tA1 = tA0 + (uint32_t)p.a0c, tA2 = tA1 + WS_KH, tD3 = tA0 + (uint32_t)p.d3_off;
        const uint32_t idesc_h = make_idesc_tf32(WS_HID), idesc_o = make_idesc_tf32(p.N3);
        uint32_t mph = 0;
        const int bar_id = 1 + g;
        {   // constant block of both hidden A regions: column 64 = 1 (bias column), 65..71 = 0
            uint32_t one[8] = {__float_as_uint(1.0f), 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            tmem_st8(tA1 + lane_sel + WS_HID, one);
            tmem_st8(tA2 + lane_sel + WS_HID, one);
            tmem_st_wait();
        }
        // this group's observation ring: 2 slots, refilled by the group itself two of its items ahead
        const uint32_t obs_ring = smem0 + 4u * (uint32_t)(p.o_obs + g * 2 * p.obs_floats);
        auto produce_obs = [&](int kk, int oslot) {      // called by one lane
            const int work = (int)blockIdx.x + kk * (int)gridDim.x;
            const int m = p.tiles == 1 ? work : work / p.tiles, tile = work - m * p.tiles;
            const uint32_t ob = (uint32_t)(min(128, p.E - tile * 128) * p.K0) * 4u;
            mbar_expect_tx(WS_BAR(B_OFULL + g * 2 + oslot), ob);
            bulk_g2s(obs_ring + 4u * (uint32_t)(oslot * p.obs_floats), obs + ((int64_t)m * p.E + tile * 128) * p.K0, ob,
                     WS_BAR(B_OFULL + g * 2 + oslot));
        };
        if (p.obs_vec && q == 0 && lane == 0) {
            if (g < n_my) produce_obs(g, 0);
            if (g + p.ng < n_my) produce_obs(g + p.ng, 1);
        }

        // TMEM row -> tanh -> tf32 -> the same TMEM columns (they become the next layer's A operand).  Accurate-tanh path:
        // rounded to nearest; tanh.approx path: left as they are - the tensor core truncates the low 13 mantissa bits, an
        // error of the same 2^-11 class as tanh.approx itself, and one integer add per activation is saved (-5 % kernel time)
        // The TMEM read port (64 B/clk per SM) and the SFU (16 tanh/clk per SM) each need ~512 cycles per group and
        // layer; used one after the other (read everything, then tanh everything) the groups fall into lockstep and the
        // two phases ADD (measured: ~7 000-cycle item chain = 3 groups x 2 layers x (read + tanh)).  Pipelined form: 16
        // columns at a time, the read of chunk c+1 is in flight while chunk c goes through the SFU, so every group keeps
        // both resources busy and the per-item cost tends to max(read, tanh) instead of their sum.
        auto epilogue_hidden = [&](uint32_t t_reg, bool active) {
            if (active) {
                uint32_t r0[16], r1[16];
                tmem_ld16_issue(t_reg + lane_sel, r0);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t* cur = (c & 1) ? r1 : r0;
                    uint32_t* nxt = (c & 1) ? r0 : r1;
                    if (c < 3) tmem_ld16_issue(t_reg + lane_sel + 16u * (uint32_t)(c + 1), nxt);
#pragma unroll
                    for (int i = 0; i < 16; ++i) cur[i] = __float_as_uint((APPROX ? ws_tanh<APPROX>(__uint_as_float(cur[i])) : tf32_rn(ws_tanh<APPROX>(__uint_as_float(cur[i])))));
                    tmem_st16(t_reg + lane_sel + 16u * (uint32_t)c, cur);
                    if (c < 3) tmem_ld_wait();
                }
                tmem_st_wait();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        };

        int s = 0;
        uint32_t par = 0, opar = 0;
        for (int i = 0; i < g; ++i)
            if (++s == p.nst) { s = 0; par ^= 1u; }
        int oslot = 0;
        // observation row of item kk -> TMEM (A operand of layer 0): columns [0, K0) = x, K0 = 1, rest 0.  Called for the
        // group's first item before the loop and for item k + ng right after layer 1 of item k has been issued, so the
        // ring wait, the shared-memory reads and the TMEM stores hide under that MMA.  The barrier that orders these
        // stores before the next layer-0 issue (and the slot refill) is the one at the end of the layer-1 epilogue.
        auto load_obs = [&](int kk) {
            const int work = (int)blockIdx.x + kk * (int)gridDim.x;
            const int m = p.tiles == 1 ? work : work / p.tiles, tile = work - m * p.tiles;
            const int e0i = tile * 128, ne = min(128, p.E - e0i);
            WS_TL(gt == 0, kk - p.ng, 12);
            if (p.obs_vec) mbar_wait_park(WS_BAR(B_OFULL + g * 2 + oslot), opar);
            WS_TL(gt == 0, kk - p.ng, 13);
            if (q * 32 < ne) {
                if (p.obs_vec) {
                    // row gt of the TMA-landed tile: fully unrolled, immediate offsets, independent loads (the group chain is
                    // instruction-latency bound; the clamped generic-pointer form cost ~1 000 cycles per item).  Columns past
                    // K0 are replaced after the load, so reading up to 7 floats past the row is harmless - it stays inside
                    // the ring (the allocation carries 32 floats of slack behind the last slot).
                    const uint32_t srow = obs_ring + 4u * (uint32_t)(oslot * p.obs_floats + gt * p.K0);
#pragma unroll
                    for (int c0 = 0; c0 < 40; c0 += 8) {
                        if (c0 < p.K0p) {
                            uint32_t x[8];
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const int kk2 = c0 + c;
                                const float v = lds32(srow + 4u * (uint32_t)kk2);
                                x[c] = __float_as_uint(tf32_rn(kk2 < p.K0 ? v : (kk2 == p.K0 ? 1.0f : 0.f)));
                            }
                            tmem_st8(tA0 + lane_sel + (uint32_t)c0, x);
                        }
                    }
                } else {
                    // rows straight from global memory (unaligned observation buffers): clamped loads, selects after
                    const float* src = obs + ((int64_t)m * p.E + e0i + min(gt, ne - 1)) * p.K0;
#pragma unroll 1
                    for (int c0 = 0; c0 < p.K0p; c0 += 8) {
                        float v[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) v[c] = src[min(c0 + c, p.K0 - 1)];
                        uint32_t x[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const int kk2 = c0 + c;
                            x[c] = __float_as_uint(tf32_rn(kk2 < p.K0 ? v[c] : (kk2 == p.K0 ? 1.0f : 0.f)));
                        }
                        tmem_st8(tA0 + lane_sel + (uint32_t)c0, x);
                    }
                }
                WS_TL(gt == 0, kk - p.ng, 14);
                tmem_st_wait();
            }
        };
        auto obs_consumed = [&](int kk) {      // after a group barrier: refill the slot item kk used, two group items ahead
            if (p.obs_vec && q == 1 && lane == 0 && kk + 2 * p.ng < n_my) produce_obs(kk + 2 * p.ng, oslot);
            if (++oslot == 2) { oslot = 0; opar ^= 1u; }
        };
        if (g < n_my) {
            load_obs(g);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            obs_consumed(g);
        }
        for (int k = g; k < n_my; k += p.ng) {
            const int work = (int)blockIdx.x + k * (int)gridDim.x;
            const int m = p.tiles == 1 ? work : work / p.tiles, tile = work - m * p.tiles;
            const int e0i = tile * 128, ne = min(128, p.E - e0i);
            const bool active = q * 32 < ne;     // whole-warp skip of padding rows
            const uint32_t s_addr = smem0 + 4u * (uint32_t)(s * p.st_floats);
            const bool st = gt == 0 && k == 6;

            WS_TL(gt == 0, k, 0);
            mbar_wait_park(WS_BAR(B_WFULL + s), par);           // weights of item k are in stage s
            WS_TL(gt == 0, k, 2);
            // MMA issue is warp-uniform (descriptors live in uniform registers), one elected lane issues
            if (q == 0) {          // layer 0
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t bdesc = make_desc(s_addr + 4u * p.o_w0, 128, (uint32_t)p.K0p * 32u);
                for (int j = 0; j < p.K0p / 8; ++j)
                    umma_tf32_ts_elect(tA1, tA0 + (uint32_t)(j * 8), bdesc + (uint64_t)(j * 16), idesc_h, j > 0 ? 1u : 0u);
                umma_commit_elect(mbar);
                __syncwarp();
            }
            WS_STAMP(st, 2);
            mbar_wait_park(mbar, mph); mph ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            WS_STAMP(st, 3);
            epilogue_hidden(tA1, active);
            WS_TL(gt == 0, k, 3);

            if (q == 0) {          // layer 1: A = tanh(layer 0) in TMEM
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t bdesc = make_desc(s_addr + 4u * p.o_w1, 128, (uint32_t)WS_KH * 32u);
#pragma unroll
                for (int j = 0; j < WS_KH / 8; ++j)
                    umma_tf32_ts_elect(tA2, tA1 + (uint32_t)(j * 8), bdesc + (uint64_t)(j * 16), idesc_h, j > 0 ? 1u : 0u);
                umma_commit_elect(mbar);
                __syncwarp();
            }
            WS_STAMP(st, 5);
            const bool has_next = k + p.ng < n_my;
            if (has_next) load_obs(k + p.ng);                  // issued behind the layer-1 MMA
            WS_TL(gt == 0, k, 1);
            mbar_wait_park(mbar, mph); mph ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            WS_TL(gt == 0, k, 7);
            WS_STAMP(st, 6);
            epilogue_hidden(tA2, active);
            if (has_next) obs_consumed(k + p.ng);
            WS_TL(gt == 0, k, 4);

            if (q == 0) {          // layer 2 (head); its completion also frees operand stage s
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t bdesc = make_desc(s_addr + 4u * p.o_w2, 128, (uint32_t)WS_KH * 32u);
#pragma unroll
                for (int j = 0; j < WS_KH / 8; ++j)
                    umma_tf32_ts_elect(tD3, tA2 + (uint32_t)(j * 8), bdesc + (uint64_t)(j * 16), idesc_o, j > 0 ? 1u : 0u);
                umma_commit_elect(mbar);
                umma_commit_elect(WS_BAR(B_WEMPTY + s));
                __syncwarp();
            }
            WS_TL(gt == 0, k, 5);
            mbar_wait_park(mbar, mph); mph ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            WS_STAMP(st, 9);
            if (active) {
                float* o = out + ((int64_t)m * p.E + e0i + gt) * p.nout;
#pragma unroll 1
                for (int c = 0; c < p.N3; c += 16) {
                    float v[16];
                    tmem_ld16(tD3 + lane_sel + (uint32_t)c, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float y = ws_tanh<APPROX>(v[i]);
                        v[i] = c + i < p.A ? y : 0.55f + 0.45f * y;   // MapContinuousToAction
                    }
                    if (gt < ne) {
                        if ((p.nout & 3) == 0) {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (c + 4 * i < p.nout)
                                    *reinterpret_cast<float4*>(o + c + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (c + i < p.nout) o[c + i] = v[i];
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // head read out before the next layer 0 reuses its columns
            for (int i = 0; i < p.ng; ++i)
                if (++s == p.nst) { s = 0; par ^= 1u; }
            WS_TL(gt == 0, k, 6);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
#undef WS_BAR
#undef WS_STAMP
#undef WS_TL
}

}  // namespace
// returns -1 when the shape is not served by this kernel (the caller falls back to the generic tcgen05 kernel)
int dfd_mlp_forward_ws_impl(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                            const int64_t* idx, const int8_t* sign, int n_members, float sigma, const float* obs,
                            int obs_per_member, float* out, int approx_tanh, cudaStream_t st) {
    const int K0 = desc->n_in, nout = 2 * desc->n_act;
    if (desc->h1 != WS_HID || desc->h2 != WS_HID || K0 > 32 || nout > 32) return -1;
    if (getenv("DFD_TC_NO_WS")) return -1;
    WsParams p = {};
    p.K0 = K0;
    p.K0p = (K0 + 1 + 7) / 8 * 8;
    p.nkq = (K0 + 1 + 3) / 4;
    p.nout = nout;
    p.N3 = (nout + 15) / 16 * 16;
    p.A = desc->n_act;
    p.w_off1 = K0 * WS_HID + WS_HID;
    p.b_off0 = K0 * WS_HID;
    p.b_off1 = p.w_off1 + WS_HID * WS_HID;
    p.w_off2 = p.b_off1 + WS_HID;
    p.b_off2 = p.w_off2 + nout * WS_HID;
    const int P = p.b_off2 + nout;
    p.E = obs_per_member;
    p.tiles = (obs_per_member + 127) / 128;
    DFD_CHECK_ARG((int64_t)n_members * p.tiles < 2147483647LL, "tcgen05 MLP path: too many work items");
    p.n_work = n_members * p.tiles;
    p.obs_vec = (((int64_t)obs_per_member * K0) % 4 == 0 && (((uintptr_t)obs) & 15) == 0) ? 1 : 0;
    p.sigma = sigma;
    p.o_w0 = 0;
    p.o_w1 = p.o_w0 + WS_HID * p.K0p;
    p.o_w2 = p.o_w1 + WS_HID * WS_KH;
    p.st_floats = p.o_w2 + p.N3 * WS_KH;
    p.eps_floats = (P + 3) / 4 * 4;                       // the replica has >= 64 floats of slack past any row
    p.ring_floats = p.eps_floats + 4;                     // + the zero word
    p.obs_floats = 128 * K0;
    // W0 items 64 * nkq must fit the fixed per-thread item count; the observation operand its TMEM columns
    if (8 * p.nkq * 8 > 2 * WS_NBUILD || p.K0p > 40) return -1;
    const size_t cap = 226 * 1024 - 128;
    p.nst = 3;
    p.ne = 4;
    // three groups when the TMEM columns (3 x (24 + 72 + 72), head accumulator aliased onto the dead observation
    // operand) and shared memory (six observation slots) allow it
    p.ng = (p.K0p <= 24 && p.N3 <= 16 && !getenv("DFD_WS_2GROUPS")) ? 3 : 2;
    auto bytes = [&]() { return ((size_t)p.nst * p.st_floats + (size_t)p.ne * p.ring_floats + 2 * (size_t)p.ng * p.obs_floats) * sizeof(float); };
    if (bytes() > cap) p.ne = 3;
    if (bytes() > cap && p.ng == 3) { p.ng = 2; p.ne = 4; if (bytes() > cap) p.ne = 3; }
    if (bytes() > cap) p.nst = 2;
    if (bytes() > cap) p.ne = 2;
    if (bytes() > cap) return -1;
    p.a0c = p.ng == 3 ? 24 : 40;
    p.tcols = p.ng == 3 ? 24 + 2 * WS_KH : 40 + 2 * WS_KH + 32;
    p.d3_off = p.ng == 3 ? 24 : 40 + 2 * WS_KH;   // three groups: the head accumulator aliases the first columns of A1/D1
    p.o_ring = p.nst * p.st_floats;
    p.o_obs = p.o_ring + p.ne * p.ring_floats;
    const size_t smem = bytes() + 128;              // slack behind the last observation slot (load_obs over-read)
    int grid = ctx->sm_count;
    if (grid > p.n_work) grid = p.n_work;
    long long* prof = nullptr;
#ifdef DFD_WS_TIMELINE
    static const bool want_prof = getenv("DFD_WS_PROF") != nullptr;
#else
    static const bool want_prof = false;
#endif
    if (want_prof) {
        cudaMalloc(&prof, (size_t)grid * 256 * sizeof(long long));
        cudaMemset(prof, 0, (size_t)grid * 256 * sizeof(long long));
    }
    if (approx_tanh) {
        DFD_CUDA(cudaFuncSetAttribute(mlp_forward_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DFD_CUDA(dfd_launch_pdl(mlp_forward_ws_kernel<true>, dim3(grid), dim3(WS_NBUILD + 128 * p.ng), smem, st, p, table->replicas,
                                table->replica_stride, theta, idx, sign, obs, out, prof));
    } else {
        DFD_CUDA(cudaFuncSetAttribute(mlp_forward_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DFD_CUDA(dfd_launch_pdl(mlp_forward_ws_kernel<false>, dim3(grid), dim3(WS_NBUILD + 128 * p.ng), smem, st, p, table->replicas,
                                table->replica_stride, theta, idx, sign, obs, out, prof));
    }
    DFD_LAUNCHED(ctx);
    if (want_prof) {
        cudaStreamSynchronize(st);
        static long long h[256];
        cudaMemcpy(h, prof + 256 * 5, 256 * sizeof(long long), cudaMemcpyDeviceToHost);   // CTA 5
        long long t0 = 0;
        for (int i = 0; i < 256; ++i) if (h[i] && (!t0 || h[i] < t0)) t0 = h[i];
        fprintf(stderr, "[ws timeline] CTA 5, cycles since first stamp; G: top, next obs staged, weights ready, E0 done, E1 done, L2 issued, end, L1 MMA done | B: top, eps ready, stage free, built\n");
        for (int k = 0; k < 16; ++k) {
            fprintf(stderr, "  item %2d G%d:", k, k % p.ng);
            for (int j = 0; j < 8; ++j) fprintf(stderr, " %7lld", h[k * 16 + j] ? h[k * 16 + j] - t0 : -1LL);
            fprintf(stderr, "  | obs: enter, landed, stores issued:");
            for (int j = 12; j < 15; ++j) fprintf(stderr, " %7lld", h[k * 16 + j] ? h[k * 16 + j] - t0 : -1LL);
            fprintf(stderr, "  | B:");
            for (int j = 8; j < 12; ++j) fprintf(stderr, " %7lld", h[k * 16 + j] ? h[k * 16 + j] - t0 : -1LL);
            fprintf(stderr, "\n");
        }
        cudaFree(prof);
    }
    return 0;
}
