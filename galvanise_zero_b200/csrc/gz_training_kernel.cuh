// gz_training_kernel.cuh — rec_unpack_kernel of gz_training.cu (sample records -> state rows, dense policy rows,
// value rows, check_sample count). In a header of its own so that tests/hostsim/warpsim_training.cpp can compile the
// same source for the CPU under a 32-thread warp emulation (the kernel only uses warp-level primitives).
#ifndef GZ_TRAINING_KERNEL_CUH
#define GZ_TRAINING_KERNEL_CUH

#include <math.h>
#include <stdint.h>

#include "../../include/gz_b200.h"

#define TRN_WARPS 8
#define TRN_MAX_ROLES 4

struct TrnArgs {
    int64_t n;
    int record_words, state_words, n_bases, n_roles, n_values, n_sym;
    int64_t n_records;
    const int32_t* records;
    const int32_t* index;          // [n] record of each output row, or null (row i = record i)
    const int32_t* sym;            // [n] symmetry of each output row, or null
    const int32_t* base_src;       // [n_sym][n_bases]: new bit j = old bit base_src[s][j]
    const int32_t* action_dest[TRN_MAX_ROLES];   // [n_sym][width_r]: new index of action a
    float* policy[TRN_MAX_ROLES];
    int width[TRN_MAX_ROLES];
    float* value;
    uint32_t* states;              // [n][state_words]
    uint32_t* prev;                // [n][state_words]
    uint8_t* prev_count;           // [n]
    int32_t* bad;                  // [1] records failing check_sample, or null
};

// one state row of the record -> out (packed the same way), through the bit permutation `src` if there is one
__device__ __forceinline__ void trn_state_row(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int sw,
                                              int n_bases, const int32_t* __restrict__ src, int lane) {
    if (!src) {
        for (int w = lane; w < sw; w += 32) {
            const int lo = w * 32;
            uint32_t v = in[w];
            if (lo + 32 > n_bases) v = lo >= n_bases ? 0u : (v & ((1u << (n_bases - lo)) - 1u));   // padding bits are zero
            out[w] = v;
        }
        return;
    }
    for (int w = 0; w < sw; ++w) {
        const int j = w * 32 + lane;
        bool bit = false;
        if (j < n_bases) {
            const int s = src[j];
            bit = (in[s >> 5] >> (s & 31)) & 1u;
        }
        const uint32_t word = __ballot_sync(0xffffffffu, bit);
        if (lane == 0) out[w] = word;
    }
}

__global__ void __launch_bounds__(TRN_WARPS * 32) rec_unpack_kernel(const TrnArgs A) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * TRN_WARPS;
    int n_bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * TRN_WARPS + (threadIdx.x >> 5); i < A.n; i += warps) {
        const int64_t ri = A.index ? (int64_t)A.index[i] : i;
        if (ri < 0 || ri >= A.n_records) {           // no such record: an all-zero row, counted as bad
            for (int w = lane; w < A.state_words; w += 32) { A.states[i * A.state_words + w] = 0u; A.prev[i * A.state_words + w] = 0u; }
            if (lane == 0) {
                A.prev_count[i] = 0;
                for (int r = 0; r < A.n_values; ++r) A.value[i * A.n_values + r] = 0.0f;
            }
            n_bad++;
            continue;
        }
        const int32_t* rec = A.records + ri * A.record_words;
        const int lead = rec[GZ_SR_LEAD], n_prev = rec[GZ_SR_N_PREV];
        int nc = rec[GZ_SR_N_CHILDREN];
        const int nc_max = (A.record_words - GZ_SR_HEADER - 2 * A.state_words) / 2;      // pairs a record can hold
        const bool nc_ok = nc >= 0 && nc <= nc_max;
        if (!nc_ok) nc = 0;
        const int s = A.sym ? A.sym[i] : 0;
        const bool sym_ok = !A.sym || (s >= 0 && s < A.n_sym);
        const int32_t* src = (A.sym && sym_ok) ? A.base_src + (int64_t)s * A.n_bases : nullptr;
        const int sw = A.state_words;
        // ---- state rows for the encoder (bases.py:242-287 reads state, then prev_states in order)
        const uint32_t* st = reinterpret_cast<const uint32_t*>(rec + GZ_SR_HEADER);
        trn_state_row(st, A.states + i * sw, sw, A.n_bases, src, lane);
        if (n_prev > 0) {
            trn_state_row(st + sw, A.prev + i * sw, sw, A.n_bases, src, lane);
        } else {
            for (int w = lane; w < sw; w += 32) A.prev[i * sw + w] = 0u;
        }
        if (lane == 0) A.prev_count[i] = n_prev > 0 ? 1 : 0;
        // ---- policy_to_array (bases.py:305-310): array[idx] = prob over the role's (legal, prob) pairs; the rows
        // were zeroed before the launch. The lead role's pairs are the root's children, every other role has its
        // single move with probability 1 (selfplaymanager.cpp:95-108).
        bool bad = !sym_ok || !nc_ok;
        const int32_t* pairs = rec + GZ_SR_HEADER + 2 * sw;
        for (int r = 0; r < A.n_roles; ++r) {
            float* row = A.policy[r] + i * A.width[r];
            const int32_t* dest = (A.sym && sym_ok) ? A.action_dest[r] + (int64_t)s * A.width[r] : nullptr;
            if (r == lead) {
                double total = 0.0;
                for (int c = lane; c < nc; c += 32) {
                    const int mv = pairs[2 * c];
                    const float p = __int_as_float(pairs[2 * c + 1]);
                    if (mv < 0 || mv >= A.width[r]) { bad = true; continue; }
                    row[dest ? dest[mv] : mv] = p;
                    if (!(-0.01 < (double)p && (double)p < 1.01)) bad = true;     // check_sample (bases.py:296-299)
                    total += (double)p;
                }
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
                if (!(0.99 < total && total < 1.01)) bad = true;                   // (bases.py:301)
            } else if (lane == 0) {
                const int mv = rec[GZ_SR_OTHER_MOVES + r];
                if (mv < 0 || mv >= A.width[r]) bad = true;
                else row[dest ? dest[mv] : mv] = 1.0f;
            }
        }
        // ---- value_to_array on the clamped final scores (datacache.py:716-718, bases.py:312-314 / :323-335)
        if (lane == 0) {
            float fs[TRN_MAX_ROLES];
            for (int r = 0; r < A.n_roles; ++r) {
                float v = __int_as_float(rec[GZ_SR_FINAL_SCORE + r]);
                v = 1.0f < v ? 1.0f : v;        // min(1.0, v)
                v = 0.0f > v ? 0.0f : v;        // max(0.0, v)
                fs[r] = v;
            }
            float* out = A.value + i * A.n_values;
            if (A.n_values == 3 && A.n_roles == 2) {
                if (!(fabs((double)fs[0] + (double)fs[1] - 1.0) < 0.01)) bad = true;
                if (fabs((double)fs[0] - 0.5) < 0.01) {
                    if (!(fabs((double)fs[1] - 0.5) < 0.01)) bad = true;
                    out[0] = 0.0f; out[1] = 0.0f; out[2] = 1.0f;
                } else {
                    out[0] = fs[0]; out[1] = fs[1]; out[2] = 0.0f;
                }
            } else {
                for (int r = 0; r < A.n_values; ++r) out[r] = r < A.n_roles ? fs[r] : 0.0f;
            }
        }
        bad = __any_sync(0xffffffffu, bad);
        if (bad) n_bad++;
    }
    if (A.bad && lane == 0 && n_bad) atomicAdd(A.bad, n_bad);
}

#endif  // GZ_TRAINING_KERNEL_CUH
