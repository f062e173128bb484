// tcgen05 / TMEM tensor-core path of the per-member MLP forward (tf32 operands, fp32 accumulate).
#include "common.cuh"

int dfd_mlp_forward_tc_impl(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                            const int64_t* idx, const int8_t* sign, int n_members, float sigma, const float* obs,
                            int obs_per_member, float* out, cudaStream_t st) {
    dfd_set_error("dfd_policy_forward: precision=1 (tcgen05) path is not built yet");
    return 4;
}
