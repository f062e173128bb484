/*
 * dfd_b200.h — C ABI of the B200-native finite-difference learner hot path.
 *
 * The reference (nexus-rl/dfd-starter) is 100 % Python and has no FFI; its
 * boundary for this path is a set of duck-typed Python objects.
 * Each entry point below names the reference code it replaces (file:line under
 * the reference root).  The host-side Python mirror of those objects lives in
 * dfd_starter_b200/*.py and binds these symbols with ctypes.
 *
 * Conventions
 *   - plain C symbols, int status return: 0 = ok, non-zero = error;
 *     dfd_last_error() returns a thread-local message for the last failure;
 *   - no C++ exceptions cross the boundary;
 *   - the CALLER owns every buffer (all pointers are device pointers unless a
 *     name says `host`); hot calls never allocate and never synchronise;
 *   - every hot call takes the CUDA stream to launch on (cudaStream_t as void*);
 *   - one context per device, not re-entrant per context;
 *   - compiled for sm_100a only; there is no CPU fallback.
 */
#ifndef DFD_B200_H
#define DFD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFD_ABI_VERSION 2

typedef struct dfd_ctx dfd_ctx;
typedef void* dfd_stream; /* cudaStream_t */

/* ---- context ----------------------------------------------------------- */
int dfd_abi_version(void);
const char* dfd_last_error(void);
int dfd_ctx_create(int device, dfd_ctx** out);
int dfd_ctx_destroy(dfd_ctx* ctx);
int dfd_ctx_sm_count(const dfd_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py's
 * gpu_launches claim is read from here, not guessed) */
int64_t dfd_ctx_launch_count(const dfd_ctx* ctx);

/* ---- noise table: utils/noise_sources.py:36-51 (SharedNoiseTable) ------- */
/* The fp32 table itself is produced on the host by numpy's legacy RandomState
 * (bit-exact requirement) and uploaded once.  On the
 * device it is kept as FOUR element-shifted replicas so that every slice
 * table[idx : idx+P] starts 16-byte aligned in replica (idx & 3):
 *     replica_s[j] = table[j + s],  s = 0..3,  row base = replica_{idx&3} + (idx & ~3)
 * plus an fp64 inclusive prefix sum of squares (prefix[i] = sum_{j<i} table[j]^2,
 * size+1 entries) so ||table[idx:idx+P]||^2 is two loads. */
typedef struct dfd_table {
    const float* replicas;   /* 4 * replica_stride floats                                   */
    int64_t replica_stride;  /* floats between replicas; multiple of 32, >= size + 64         */
    const double* prefix_sq; /* size + 1 doubles                                             */
    int64_t size;            /* number of table entries                                      */
} dfd_table;

int64_t dfd_table_replica_stride(int64_t size);
/* table_dev: `size` floats already on the device.  Fills replicas (4*stride
 * floats) and prefix_sq (size+1 doubles); scratch needs dfd_table_scratch_bytes(size). */
size_t dfd_table_scratch_bytes(int64_t size);
int dfd_table_build(dfd_ctx* ctx, const float* table_dev, int64_t size, float* replicas, int64_t replica_stride,
                    double* prefix_sq, void* scratch, size_t scratch_bytes, dfd_stream stream);

/* sigma-scaled fp16 mirror of the table for the "direct-from-table" forward of wide MuJoCo MLPs
 * (csrc/mlp_forward_direct.cu): a Linear layer is linear in its weights, x.(theta + s*sigma*eps)^T =
 * x.theta^T + s*(x.(sigma*eps)^T), so a member's weight tiles are never built - both terms are tcgen05 MMAs whose B
 * operands come by TMA from (i) an fp16 copy of theta and (ii) this mirror: eight element-shifted replicas of
 * fp16(fl32(sigma*table[j])) (the first rounding is worker.py:28's own) so that any table[idx+off:...] slice starts
 * 16-byte aligned.  buf: CALLER-owned device memory of dfd_table_scaled16_bytes(size, n_params) bytes, 256-byte aligned,
 * which also carries the fp16 theta scratch for policies of up to n_params parameters; it must stay alive until
 * dfd_table_drop_scaled16 / dfd_ctx_destroy.  One mirror per context (one table, one sigma); dfd_policy_forward uses it
 * when `table` and `sigma` match and dfd_policy_direct_supported(desc), and the streaming kernel otherwise. */
size_t dfd_table_scaled16_bytes(int64_t size, int64_t n_params);
int dfd_table_build_scaled16(dfd_ctx* ctx, const dfd_table* table, float sigma, int64_t n_params, void* buf, size_t bytes,
                             dfd_stream stream);
int dfd_table_drop_scaled16(dfd_ctx* ctx);

/* ---- perturbation: worker/worker.py:28  new_flat = flat + sigma * eps ---- */
/* out[m, :] = theta + sign[m]*sigma*table[idx[m] : idx[m]+P], fp32, product
 * rounded then sum rounded (no FMA) so the result is bit-identical to numpy.
 * Materialising members is for parity tests and generic host policies; the
 * forward kernels below generate the same values in-kernel and never write
 * them to HBM.  out_stride in floats (>= P). */
int dfd_perturb_members(dfd_ctx* ctx, const dfd_table* table, const float* theta, int64_t n_params,
                        const int64_t* idx, const int8_t* sign, int n_members, float sigma, float* out,
                        int64_t out_stride, dfd_stream stream);

/* ---- policy forwards: policies/policy.py:26-29 + mujoco/discrete/atari/impala */
enum { DFD_POLICY_MUJOCO = 0, DFD_POLICY_DISCRETE = 1, DFD_POLICY_ATARI = 2, DFD_POLICY_IMPALA = 3 };

typedef struct dfd_policy_desc {
    int kind;      /* DFD_POLICY_*                                                         */
    int n_in;      /* MLPs: observation width K                                            */
    int h1, h2;    /* MLPs: hidden widths (reference: 64, 64; mujoco.py:33-34)             */
    int n_act;     /* actions A (MuJoCo head emits 2A: mean | std)                         */
    int precision; /* 0 = fp32 CUDA cores (exact path);
                      1 = tensor cores.  MuJoCo 64x64 nets: tf32 tcgen05, weights resident; wide MuJoCo nets and Atari:
                          fp16-operand tcgen05 with TMA-fed weight tiles when the scaled table mirror is registered
                          (dfd_table_build_scaled16), else tf32 tcgen05 with weights built in shared memory (MuJoCo) /
                          the exact kernel (Atari); IMPALA: mma.sync convolutions with fp16 operands, fp32 dense tail;
                      2 = MuJoCo: as 1 + single-instruction tanh.approx (2^-11 relative); IMPALA: as 1 with the dense
                          tail (Linear 2048->256 + LSTM, 90 % of the parameters) as TMA-fed tcgen05 GEMMs when the
                          scaled mirror is registered (csrc/impala_tail.cuh);
                      3 = IMPALA: tcgen05 trunk as well (implicit GEMMs over fp16 operand PLANES in the UMMA no-swizzle
                          layout - a filter tap is a shifted descriptor start address, no im2col -, the residual stream
                          in TMEM; csrc/impala_forward_tc.cu): the fastest level (C5 forward 587 us against 762).
                      fp16 operands carry tf32's 10-bit mantissa; accumulation is fp32 everywhere.  */
} dfd_policy_desc;

int64_t dfd_policy_num_params(const dfd_policy_desc* desc);
/* 1 when the direct-from-table tensor path serves this MuJoCo shape (n_in % 8 == 0, hidden widths 128 or 256, 2A <= 48) */
int dfd_policy_direct_supported(const dfd_policy_desc* desc);
int64_t dfd_policy_num_buffers(const dfd_policy_desc* desc);
int64_t dfd_policy_out_width(const dfd_policy_desc* desc);

/* Batched member x observation forward (the new surface of SURVEY.md §8b):
 *   member m uses theta + sign[m]*sigma*table[idx[m]:idx[m]+P] (sign 0 = unperturbed
 *   eval member, worker.py:23-25,35); obs is [n_members, obs_per_member, obs_width];
 *   out is [n_members, obs_per_member, out_width]:
 *     MuJoCo   mean(A) | std(A)            (mujoco.py:35-41, torch_helpers.py:20-25)
 *     Discrete probs(A)                    (discrete.py:37-48)
 *     Atari    probs(A), obs NCHW 4x84x84  (atari.py:35-51)
 *   bn_buffers: BatchNorm running stats in state_dict order (mean, var, num_batches_tracked
 *   per BN layer), shared by all members; NULL for MuJoCo. */
int dfd_policy_forward(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                       const float* bn_buffers, const int64_t* idx, const int8_t* sign, int n_members, float sigma,
                       const float* obs, int obs_per_member, float* out, dfd_stream stream);

/* IMPALA CNN+LSTM (impala.py:136-186): E independent single-step environments
 * per member, each with its own carried (h, c) of 256 floats.
 * frame [M,E,3,64,64] (0..255), reward [M,E], done [M,E] (uint8),
 * h_in/c_in/h_out/c_out [M,E,256], probs [M,E,A]. scratch: dfd_impala_scratch_bytes.
 * With an even member count, members j and j + M/2 are evaluated by one CTA; when they
 * share their table index (an antithetic pair in [plus | minus] order) theta and the eps
 * row of the dense tail are streamed once for both. */
size_t dfd_impala_scratch_bytes(int n_members, int obs_per_member);
int dfd_impala_forward(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                       const float* bn_buffers, const int64_t* idx, const int8_t* sign, int n_members, float sigma,
                       const float* frame, const float* reward, const uint8_t* done, const float* h_in,
                       const float* c_in, int obs_per_member, float* probs, float* h_out, float* c_out,
                       void* scratch, size_t scratch_bytes, dfd_stream stream);

/* ---- the estimator: learner/finite_differences.py:24-114 ---------------- */
/* One learner step on the device, in three hot calls.
 *
 * (1) dfd_fd_prepare  — finite_differences.py:40-43 (baseline, standardise),
 *     :87-89,107 (lambda_i = sign_i*sigma*eps_i + d_{e_i}, ||lambda_i||^2).
 *     Inputs are the ACCEPTED returns only (the host applies the epoch test of
 *     :82-85): reward[n] (fp64), idx[n], sign[n] (+1/-1), hist_row[n] (-1 = return
 *     from the current epoch, dist = 0; else row of `dist` [n_hist, dist_stride]
 *     holding theta_e - theta_now).  For delayed rows the eps_i . d_e dot products
 *     are computed by this call (a first pass over those rows only).
 *     If `paired` != 0 the n returns are [R plus-members | R minus-members] of the
 *     same R table rows and their coefficients are merged so each row is read once.
 *     If stats_reward != NULL the mean / std are taken over stats_reward[n_stats]
 *     (multi-GPU: every rank standardises its shard with the statistics of ALL
 *     ranks' returns); otherwise over reward[n_returns].
 *     Outputs the row list of step (2): row_ptr[n_rows], row_coef[n_rows] with
 *     n_rows = (paired ? n_returns/2 : n_returns) + n_hist  (known to the host
 *     without a synchronisation).
 * (2) dfd_fd_reduce   — finite_differences.py:49  g = sum_i w_i * lambda_i/||lambda_i||^2
 *     as ONE streaming pass:  g[p] = sum_r row_coef[r] * row_ptr[r][p].
 * (3) dfd_dsgd_step   — dsgd/dynamic_sgd.py:18-39 + finite_differences.py:54-78:
 *     theta += lr*sqrt(P)*lr_scale * g/||g|| (the learner hands -g to DSGD, which
 *     subtracts), update_size = ||theta_old - theta_new||, then the theta-history
 *     ring and the dist rows (theta_e - theta_new) are refreshed in place.
 */
typedef struct dfd_fd_rows {
    const float** row_ptr; /* [max_rows] device array of 16-byte aligned row base pointers */
    float* row_coef;       /* [max_rows]                                                  */
    int max_rows;
} dfd_fd_rows;

size_t dfd_fd_prepare_scratch_bytes(int n_returns, int n_hist);
int dfd_fd_prepare(dfd_ctx* ctx, const dfd_table* table, int64_t n_params, const double* reward, const int64_t* idx,
                   const int8_t* sign, const int32_t* hist_row, int n_returns, int paired, double baseline,
                   float sigma, const float* dist, int64_t dist_stride, int n_hist, const double* stats_reward,
                   int n_stats, dfd_fd_rows* rows, void* scratch, size_t scratch_bytes, dfd_stream stream);

size_t dfd_fd_reduce_scratch_bytes(const dfd_ctx* ctx, int64_t n_params, int n_rows);
int dfd_fd_reduce(dfd_ctx* ctx, const dfd_fd_rows* rows, int n_rows, int64_t n_params, float* grad, void* scratch,
                  size_t scratch_bytes, dfd_stream stream);

/* hist: [hist_cap, hist_stride] ring of past theta (policy_history), dist: same
 * shape.  For the n_hist_valid rows that hold data BEFORE this call,
 * dist[r] = hist[r] - theta_new (finite_differences.py:66-73); then theta_new is
 * written to ring slot hist_write_row (-1: do not record; :75-78).
 * update_size_out: one device float.  All scratch buffers of the fd_* / dsgd
 * calls must be zero-filled once when allocated (their counters self-reset). */
int dfd_dsgd_step(dfd_ctx* ctx, float* theta, const float* grad, int64_t n_params, double lr, double lr_scale,
                  float* hist, float* dist, int64_t hist_stride, int n_hist_valid, int hist_write_row,
                  float* update_size_out, void* scratch, size_t scratch_bytes, dfd_stream stream);
size_t dfd_dsgd_scratch_bytes(int64_t n_params);

/* ---- population sharded over GPUs: the one exchange step --------------------------
 * One process per GPU.  Every rank evaluates its slice of the antithetic pairs and runs
 *   dfd_fd_prepare_partial -> dfd_fd_reduce -> dfd_xchg_allreduce -> dfd_dsgd_step.
 * dfd_fd_prepare_partial: dfd_fd_prepare for current-epoch antithetic pairs with the standardisation of
 *   finite_differences.py:43 DEFERRED: coefficients are (R+ - R-) * sigma / ||sigma*eps||^2 (the mean cancels
 *   inside a pair, 1/std is common to all rows) and this rank's (sum x, sum x^2, n, min, max) of
 *   x = reward - baseline are written to stats_out[5] (device doubles).  hist_row must be all -1.
 * dfd_xchg_allreduce: ONE kernel pushes the partial gradient and the statistics into every peer's mailbox
 *   over NVLink (peer stores into CUDA-IPC mapped memory), waits for all peers, sums the world partials in
 *   rank order (bitwise identical on every rank), applies 1/std(all rewards) (identity when all rewards are
 *   equal, utils/math_helpers.py:131-133) and writes grad_out.  No host synchronisation, CUDA-graph
 *   replayable; every rank must call it the same number of times.
 * Mailboxes: each rank creates one (dfd_xchg_mailbox_create, zero-filled, returns the 64-byte CUDA IPC
 *   handle), the handles are exchanged by the host (torch.distributed all_gather_object in dist.py), every
 *   rank opens its peers' (dfd_xchg_mailbox_open) and passes a DEVICE array of world pointers with entry r =
 *   rank r's mailbox as mapped in this process (entry `rank` = its own). */
int dfd_fd_prepare_partial(dfd_ctx* ctx, const dfd_table* table, int64_t n_params, const double* reward,
                           const int64_t* idx, const int8_t* sign, const int32_t* hist_row, int n_returns,
                           double baseline, float sigma, dfd_fd_rows* rows, double* stats_out, void* scratch,
                           size_t scratch_bytes, dfd_stream stream);
size_t dfd_xchg_mailbox_bytes(int64_t n_params, int world);
int dfd_xchg_mailbox_create(dfd_ctx* ctx, size_t bytes, void** mailbox, unsigned char* ipc_handle64);
int dfd_xchg_mailbox_open(dfd_ctx* ctx, const unsigned char* ipc_handle64, void** peer_mailbox);
int dfd_xchg_mailbox_close(dfd_ctx* ctx, void* peer_mailbox);
int dfd_xchg_mailbox_destroy(dfd_ctx* ctx, void* mailbox);
int dfd_xchg_allreduce(dfd_ctx* ctx, void* const* mailboxes, int rank, int world, int64_t n_params,
                       const float* grad_partial, const double* stats5, float* grad_out, dfd_stream stream);
/* All-gather of n doubles per rank (every rank the SAME n) over the same mailboxes and step counter: dst[world][n] in rank
 * order.  The sharded learner's fd_state batches (returns from older epochs: per-return norms are rank-local, the
 * standardisation of finite_differences.py:40-43 is not deferrable) gather the rewards with this call, form the
 * coefficients with the statistics of ALL ranks' returns (dfd_fd_prepare, stats_reward = dst), reduce, and sum the
 * partial gradients with dfd_xchg_allreduce(stats5 = five zeros: count 0 means "already standardised").  n * 8 + 256 bytes
 * must fit a slot of the n_params-sized mailbox; n == 0 on every rank is a no-op.  No NCCL call on the step. */
int dfd_xchg_gather_f64(dfd_ctx* ctx, void* const* mailboxes, int rank, int world, int64_t n_params, const double* src,
                        int n, double* dst, dfd_stream stream);

/* ---- the whole step after the returns are known, as ONE kernel (short parameter vectors) ---------------
 * dfd_fd_step_fused = dfd_fd_prepare + dfd_fd_reduce [+ dfd_xchg_allreduce] + dfd_dsgd_step for fd_return-mode
 * batches (all returns from the current epoch) with n_params <= 32 768: learner/finite_differences.py:40-49,
 * 54-78 and dsgd/dynamic_sgd.py:18-39 in one grid of co-resident CTAs (csrc/fd_tail.cu).  world == 1: single
 * GPU, mailboxes may be NULL.  world > 1: antithetic pairs only; mailboxes as for dfd_xchg_allreduce (the two
 * entry points share the mailbox layout and step counter, a learner may use either on any step).
 * dfd_fd_step_fused_scratch_bytes returns 0 when the shape is not served (the caller then uses the separate
 * calls); scratch must be zero-filled once when allocated. */
size_t dfd_fd_step_fused_scratch_bytes(const dfd_ctx* ctx, int64_t n_params, int n_returns, int paired);
int dfd_fd_step_fused(dfd_ctx* ctx, const dfd_table* table, int64_t n_params, const double* reward, const int64_t* idx,
                      const int8_t* sign, int n_returns, int paired, double baseline, float sigma, float* theta,
                      float* grad, double lr, double lr_scale, float* hist, float* dist, int64_t hist_stride,
                      int n_hist_valid, int hist_write_row, float* update_size_out, void* const* mailboxes, int rank,
                      int world, void* scratch, size_t scratch_bytes, dfd_stream stream);

/* ---- observation normalisation and per-member observation statistics -----------------
 * dfd_normalize_obs: out = clip((obs - mean) / std, -clip, clip), worker/agent.py:40-41 applied to a whole batch of
 * observations [n_rows, width] (in place allowed); mean / std: [width] device values (the learner-wide
 * WelfordRunningStat.mean / .std, utils/math_helpers.py:46-66), float when stats_f64 == 0, double otherwise.
 * stats_f64 == 0: fp32 IEEE subtract and divide, bit-identical to numpy on fp32 inputs.  stats_f64 != 0: the
 * reference WORKER's arithmetic - its statistics were deserialised from FDState.obs_stats as float64 arrays
 * (worker/worker.py:43, math_helpers.py:92-101), so it normalises in fp64 and rounds to fp32 once
 * (policies/policy.py:28); bit-identical to that.
 * dfd_member_obs_stats: the statistics every member's agent accumulates over its own observations
 * (agent.py:38-39 -> WelfordRunningStat.update, math_helpers.py:29-39): obs [n_members, obs_per_member, width],
 * select [n_members, obs_per_member] (non-zero = this observation was drawn for the update), out
 * [n_members, 2*width + 1] = running_mean | running_variance | count, the layout of `serialize()` (:89-90) that
 * travels as FDReturn.obs_stats_update.  Sequential fp32 in the reference's order: bit-identical. */
int dfd_normalize_obs(dfd_ctx* ctx, const float* obs, int64_t n_rows, int width, const void* mean, const void* stdv,
                      int stats_f64, float clip, float* out, dfd_stream stream);
int dfd_member_obs_stats(dfd_ctx* ctx, const float* obs, const uint8_t* select, int n_members, int obs_per_member,
                         int width, float* out, dfd_stream stream);

/* ---- small host <-> device staging without the copy engine ------------------------------------------------
 * The per-step small transfers of this path - the batch arrays going up (the FDReturn fields the estimator reads,
 * learner/finite_differences.py:94-114) and the results coming back (rewards, `update_size`, theta for
 * `get_trainable_flat`, finite_differences.py:54-59) - are moved by a kernel that accesses the PINNED host buffer
 * through its device alias, so they neither queue on a DMA copy engine behind a large observation upload nor block
 * the host.  src / dst: each either device memory or page-locked host memory (cudaHostAlloc / torch pin_memory),
 * both 16-byte aligned.  Stream-ordered like cudaMemcpyAsync: the host may touch a pinned source again / read a
 * pinned destination once the stream has passed this call. */
int dfd_host_stage(dfd_ctx* ctx, const void* src, void* dst, size_t bytes, dfd_stream stream);

/* ---- return ingestion from the RPC loop (host only) -------------------------------
 * Replaces the per-return object path networking/server.py:151-162 (SubmitReturn / SubmitReturns ->
 * FDReturn.deserialize_from_grpc, learner/fd_return.py:41-56) feeding learner/finite_differences.py:94-114:
 * the serialized proto3 `Return` / `ReturnArray` bytes
 * (networking/rpc_misc/proto/client_server_interface.proto:30-47) are decoded straight into the
 * structure-of-arrays batch the device learner uploads.  All pointers are HOST pointers owned by the caller; no
 * CUDA call is made.  `encoded_noise` keys that are a decimal table index (optionally '+' / '-' prefixed, the
 * antithetic extension) are parsed into idx / sign; any other key leaves idx = -1, sign = 0 and is found at
 * buf[key_off .. key_off + key_len).  eval_states / eval_states_shape / obs_stats_update are returned as the byte
 * range of their packed payload inside buf (little-endian fp32 / varints), length 0 when absent.
 * dfd_wire_count_returns: number of `rets` in a ReturnArray.  dfd_wire_decode_returns: number of returns decoded
 * (is_array = 0: buf is ONE Return).  Negative results: DFD_WIRE_MALFORMED, or DFD_WIRE_UNSUPPORTED for the legal
 * but unpacked / split repeated-field encodings no proto3 writer of the reference produces (the caller then takes
 * its general decoder). */
#define DFD_WIRE_MALFORMED (-1)
#define DFD_WIRE_UNSUPPORTED (-2)
typedef struct dfd_return_soa {
    int64_t* epoch;
    int64_t* idx;
    int8_t* sign;
    double* reward;
    float* novelty;
    float* entropy;
    int32_t* timesteps;
    uint8_t* is_eval;
    int64_t* key_off;
    int32_t* key_len;
    int64_t* states_off;
    int32_t* states_len;
    int64_t* shape_off;
    int32_t* shape_len;
    int64_t* stats_off;
    int32_t* stats_len;
} dfd_return_soa;
int64_t dfd_wire_count_returns(const uint8_t* host_buf, size_t len);
int64_t dfd_wire_decode_returns(const uint8_t* host_buf, size_t len, int is_array, int64_t max_returns,
                                const dfd_return_soa* out);

/* ---- strategy distances / novelty of many members --------------------------------
 * A strategy is a policy head evaluated on the zeta frames: [n_frames, width] floats (`get_strategy`,
 * policies/mujoco.py:29-30, discrete.py:31-32; produced here by dfd_policy_forward with the zeta frames as
 * observations).  dist(a_i, b_j) follows utils/math_helpers.py:166-222 (the per-frame term in fp32, in the
 * reference's operation order; the mean over frames in fp64):
 *   DFD_DIST_L2                         l2_dist                                    :166-170
 *   DFD_DIST_CATEGORICAL_TVD            categorical_tvd                            :218-221
 *   DFD_DIST_GAUSSIAN_WASSERSTEIN       gaussian_wasserstein_dist_from_strategies  :200-216 (mean | std halves)
 *   DFD_DIST_CATEGORICAL_BHATTACHARYYA  categorical_bhattacharrya_dist             :194-197
 *   DFD_DIST_GAUSSIAN_BHATTACHARYYA     gaussian_bhattacharrya_dist                :173-191 (as written there)
 * a: [n_a, n_frames, width], b: [n_b, n_frames, width] (device).  dists (nullable): [n_a, n_b] doubles.
 * row_min (nullable): [n_a] doubles = min_j dist(a_i, b_j), i.e. `compute_strategy_novelty` (:147-155,
 * strategy/strategy_handler.py:25-30) of every member against the history in one launch; +inf when n_b == 0.
 * exclude_diagonal != 0 leaves j == i out of row_min (nearest OTHER point when a and b are the same set:
 * strategy/sparse_history_manager.py:48-70, 111-149). */
#define DFD_DIST_L2 0
#define DFD_DIST_CATEGORICAL_TVD 1
#define DFD_DIST_GAUSSIAN_WASSERSTEIN 2
#define DFD_DIST_CATEGORICAL_BHATTACHARYYA 3
#define DFD_DIST_GAUSSIAN_BHATTACHARYYA 4
int dfd_strategy_distances(dfd_ctx* ctx, const float* a, int n_a, const float* b, int n_b, int n_frames, int width,
                           int kind, double* dists, double* row_min, int exclude_diagonal, dfd_stream stream);

/* ---- RNGNoiseSource drawn on the device: utils/noise_sources.py:4-20 ------ */
/* The reference's default noise source (run_sequential.py:89, run_server.py:78, run_client.py:123): the key of a
 * member is the "state,inc" of a numpy PCG64 and its noise is Generator.standard_normal(n_params) drawn from that state
 * (sample :10-13 on the worker, decode :15-20 once per return on the learner, finite_differences.py:87).  This call
 * draws the rows of a whole batch, bit-identical to numpy's 256-layer ziggurat over PCG64 (csrc/rng_normal_core.h has
 * the algorithm and how parity is pinned):
 *   streams       device, n_streams x {state_lo, state_hi, inc_lo, inc_hi}: the PCG64 state each stream starts from;
 *   every stream yields rows_per_stream consecutive rows of n_params normals (learner: one stream per return,
 *                 rows_per_stream = 1; worker: ONE stream whose rows are its consecutive sample() calls);
 *   row r of the launch (stream * rows_per_stream + row) goes to row dest_row[r] (device, nullable = r) of
 *                 rows_out (fp32, nullable) and / or rows_out_f64 (the raw normals, nullable), row_stride elements apart:
 *                 theta == NULL: rows_out = fp32(eps)                       (np.asarray(decode(key), float32))
 *                 theta != NULL: rows_out = fp32(fp64(theta) + sigma * eps) (worker.py:28 + policy.py:40-42: product
 *                                and sum rounded separately in fp64, then the fp32 cast of set_trainable_flat);
 *   row_words     device, n_streams x (rows_per_stream + 1) int64: 64-bit words the stream has consumed when row r
 *                 begins (entry rows_per_stream: when the last row ends) - the host advances the key by that count to
 *                 name the next row's key and to leave its generator where numpy would have left it;
 *   row_state     device, nullable, n_streams x (rows_per_stream + 1) x {lo, hi}: the PCG64 state at the same points
 *                 (the "state" half of row r's key, ready to print);
 *   status        device, one word, 0 = done.  DFD_RNG_SHORT: the word budget (margin x normals + 1024; < 1 selects
 *                 1.04, the mean is 1.022) ended early - call again with a larger margin; DFD_RNG_TAILCAP: a tail loop
 *                 ran past 60 rounds (probability < 1e-60; reported, not followed); DFD_RNG_SERIAL (informational): a
 *                 stream's chunk entries were resolved by the serial pass.  There is no "approximately equal" case:
 *                 the two libm calls of numpy's ziggurat (log1p, exp) are restated operation by operation;
 *   libm_fused    which of glibc's two builds of log1p / exp the host's numpy calls (1: the -mfma builds that x86-64
 *                 CPUs with FMA + AVX2 select; the host probes a few arguments on which the builds differ);
 *   force_serial  != 0 resolves every stream serially (tests);
 *   scratch       256-byte aligned, dfd_rng_scratch_bytes(n_streams, rows_per_stream, n_params, margin) bytes. */
#define DFD_RNG_TAILCAP 2u
#define DFD_RNG_SHORT 4u
#define DFD_RNG_SERIAL 8u
size_t dfd_rng_scratch_bytes(int n_streams, int64_t rows_per_stream, int64_t n_params, double margin);
int dfd_rng_normal_rows(dfd_ctx* ctx, const uint64_t* streams, int n_streams, int64_t rows_per_stream, int64_t n_params,
                        const float* theta, double sigma, const int32_t* dest_row, float* rows_out, double* rows_out_f64,
                        int64_t row_stride, int64_t* row_words, uint64_t* row_state, uint32_t* status, int libm_fused,
                        int force_serial,
                        double margin, void* scratch, size_t scratch_bytes, dfd_stream stream);

/* ---- synthetic return (bench / tests only) ------------------------------- */
/* Stand-in for the environment, which is outside this path (worker/agent.py is
 * out of scope): reward[m] = -mean_{e,j}(out[m,e,j]-target[j])^2 (fp64). */
int dfd_synthetic_reward(dfd_ctx* ctx, const float* out, int n_members, int obs_per_member, int out_width,
                         const float* target, double* reward, dfd_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* DFD_B200_H */
