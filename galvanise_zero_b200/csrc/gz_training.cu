// gz_training.cu — sample records -> training arrays, batched on the device (sm_100a).
//
// The step after self-play in the reference is the trainer's DataCache.sync
// (src/ggpzero/nn/datacache.py:673-760): for every sample of a generation file, optionally translated by a board
// symmetry (augment_data, :597-671), it clamps final_score to [0, 1] (:716-718) and builds one row per network
// input / output with the transformer (src/ggpzero/nn/bases.py): state_to_channels(state, prev_states) (:242-287),
// policy_to_array(policy, role) (:305-310) for every role and value_to_array(final_score) (:312-314, draw head:
// GdlBasesTransformer_Draws :323-335); check_sample (:289-303) validates the record. Here the same arrays are
// produced for a whole batch of the engine's fixed-size sample records (gz_puct_core.h SR_*; what
// gz_puct_fetch_samples returns and distributed.gather_sample_records gathers) without leaving the device:
//
//   rec_unpack_kernel   one warp per record: state / previous state rows (through the symmetry's bit permutation
//                       when the record has one: a lane per bit, one ballot per word), the sparse policy pairs
//                       scattered into the zeroed dense rows (action indices through the symmetry's action map),
//                       the value row, the check_sample predicates
//   encode kernels      (gz_encoder.cu) over the unpacked state rows -> planes
//
// HBM-bound byte work: per record it reads the record once and writes the planes and the dense policy rows.
#include <cuda_runtime.h>
#include <stdint.h>

#include "gz_common.cuh"
#include "gz_transformer.h"

extern "C" int gz_encode_layout(gz_transformer* t, int64_t n, int state_words, const uint32_t* d_states,
                                const uint32_t* d_prev, const uint8_t* d_prev_count, int max_prev, void* d_out,
                                int out_dtype, int c_pad, void* stream);

#include "gz_training_kernel.cuh"

extern "C" int64_t gz_records_to_training_scratch_bytes(int64_t n, int state_words) {
    if (n < 0 || state_words < 1) return -1;
    const int64_t row = (int64_t)state_words * 4;
    auto up = [](int64_t b) { return (b + 255) & ~(int64_t)255; };
    return up(n * row) + up(n * row) + up(n);
}

extern "C" int gz_records_to_training(gz_transformer* t, int64_t n, int64_t n_records, int record_words,
                                      int state_words, int n_roles, const int32_t* d_records,
                                      const int32_t* d_index, const int32_t* d_sym, int n_sym,
                                      const int32_t* d_base_src, const int32_t* const* d_action_dest, void* d_planes,
                                      int plane_dtype, float* const* d_policy, float* d_value, int n_values,
                                      int32_t* d_bad, void* d_scratch, void* stream) {
    if (!t || !t->finalized) { gz_set_error("transformer is null or not finalised"); return GZ_ERR_ARG; }
    if (n_records < 0 || (!d_index && n_records < n)) { gz_set_error("fewer records than output rows and no row index"); return GZ_ERR_ARG; }
    if (n < 0 || n_roles < 1 || n_roles > TRN_MAX_ROLES || n_roles != (int)t->policy_sizes.size()) {
        gz_set_error("bad batch size or role count (transformer has %d policy heads)", (int)t->policy_sizes.size());
        return GZ_ERR_ARG;
    }
    if (state_words * 32 < t->n_bases || record_words < GZ_SR_HEADER + 2 * state_words) {
        gz_set_error("record_words / state_words do not hold a state of %d bases", t->n_bases);
        return GZ_ERR_ARG;
    }
    if (n_values != t->num_rewards) { gz_set_error("n_values must be the transformer's num_rewards (%d)", t->num_rewards); return GZ_ERR_ARG; }
    if (n_values != n_roles && !(n_values == 3 && n_roles == 2)) {
        gz_set_error("value rows: one per role, or the draw head (3 values for 2 roles)");
        return GZ_ERR_ARG;
    }
    if (t->num_prev_states > 1) { gz_set_error("sample records carry at most one previous state"); return GZ_ERR_ARG; }
    if (n == 0) return GZ_OK;
    if (!d_records || !d_planes || !d_policy || !d_value || !d_scratch) { gz_set_error("null device pointer"); return GZ_ERR_ARG; }
    if (d_sym && (n_sym < 1 || !d_base_src || !d_action_dest)) { gz_set_error("symmetry ids need their tables"); return GZ_ERR_ARG; }
    GZ_CUDA(cudaSetDevice(t->device));
    cudaStream_t s = (cudaStream_t)stream;
    TrnArgs A{};
    A.n = n;
    A.record_words = record_words;
    A.state_words = state_words;
    A.n_bases = t->n_bases;
    A.n_roles = n_roles;
    A.n_values = n_values;
    A.n_sym = d_sym ? n_sym : 0;
    A.n_records = n_records;
    A.records = d_records;
    A.index = d_index;
    A.sym = d_sym;
    A.base_src = d_base_src;
    for (int r = 0; r < n_roles; ++r) {
        if (!d_policy[r] || (d_sym && !d_action_dest[r])) { gz_set_error("null policy / action table of role %d", r); return GZ_ERR_ARG; }
        A.policy[r] = d_policy[r];
        A.width[r] = t->policy_sizes[r];
        A.action_dest[r] = d_sym ? d_action_dest[r] : nullptr;
        if (record_words < GZ_SR_HEADER + 2 * state_words + 2) { gz_set_error("record too short"); return GZ_ERR_ARG; }
        GZ_CUDA(cudaMemsetAsync(d_policy[r], 0, (size_t)n * A.width[r] * sizeof(float), s));
    }
    A.value = d_value;
    const int64_t row = (int64_t)state_words * 4;
    auto up = [](int64_t b) { return (b + 255) & ~(int64_t)255; };
    uint8_t* base = (uint8_t*)d_scratch;
    A.states = (uint32_t*)base;
    A.prev = (uint32_t*)(base + up(n * row));
    A.prev_count = base + 2 * up(n * row);
    A.bad = d_bad;
    if (d_bad) GZ_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int32_t), s));
    const int64_t blocks = (n + TRN_WARPS - 1) / TRN_WARPS;
    const int grid = (int)(blocks < (int64_t)t->n_sms * 8 ? blocks : (int64_t)t->n_sms * 8);
    rec_unpack_kernel<<<grid, TRN_WARPS * 32, 0, s>>>(A);
    GZ_CUDA(cudaGetLastError());
    const bool with_prev = t->num_prev_states > 0;
    return gz_encode_layout(t, n, state_words, A.states, with_prev ? A.prev : nullptr,
                            with_prev ? A.prev_count : nullptr, with_prev ? 1 : 0, d_planes, plane_dtype, 0, stream);
}
