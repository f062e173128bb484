// Atari / IMPALA perturbed forwards (policies/atari.py:35-51, policies/impala.py:136-186).
#include "common.cuh"

int dfd_atari_forward_impl(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                           const float* bn_buffers, const int64_t* idx, const int8_t* sign, int n_members, float sigma,
                           const float* obs, int obs_per_member, float* out, cudaStream_t st) {
    dfd_set_error("dfd_policy_forward: the Atari forward is not built yet");
    return 4;
}

extern "C" size_t dfd_impala_scratch_bytes(int n_members, int obs_per_member) { return 256; }

extern "C" int dfd_impala_forward(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                                  const float* bn_buffers, const int64_t* idx, const int8_t* sign, int n_members,
                                  float sigma, const float* frame, const float* reward, const uint8_t* done,
                                  const float* h_in, const float* c_in, int obs_per_member, float* probs, float* h_out,
                                  float* c_out, void* scratch, size_t scratch_bytes, dfd_stream stream) {
    dfd_set_error("dfd_impala_forward: not built yet");
    return 4;
}
