// gz_encoder.cu — base states -> NCHW input planes (bf16 or f32), batched.
//
// Replaces GdlBasesTransformer::toChannels (src/cpp/gdltransformer.cpp:22-52) and the builder
// methods addBoardBase / addControlBase (src/cpp/gdltransformer.h:87-95, exported to Python at
// src/cpp/pyobjects/gdltransformer_impl.cpp:27-48,186-219). The kernel is output-oriented: every
// thread produces 8 consecutive plane elements and stores them with one 128-bit (bf16) or two
// 128-bit (f32) writes; the scatter tables of the reference are inverted at finalise time.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/gz_b200.h"
#include "gz_common.cuh"
#include "gz_encoder.cuh"
#include "gz_transformer.h"

#define ENC_STATES_PER_BLOCK 8
#define ENC_THREADS 256

template <typename OutT>
__device__ __forceinline__ OutT enc_cast(float v);
template <>
__device__ __forceinline__ float enc_cast<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 enc_cast<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename OutT>
__global__ void __launch_bounds__(ENC_THREADS)
encode_kernel(const __grid_constant__ EncParams E, int64_t n, int sw, const uint32_t* __restrict__ states,
              const uint32_t* __restrict__ prev, const uint8_t* __restrict__ prev_count, int max_prev,
              OutT* __restrict__ out, int c_pad) {
    extern __shared__ __align__(16) uint32_t esm[];
    uint32_t* cur = esm;                                      // [SB][sw]
    uint32_t* por = cur + ENC_STATES_PER_BLOCK * sw;          // [SB][sw]  OR of previous states
    float* ctrl = reinterpret_cast<float*>(por + ENC_STATES_PER_BLOCK * sw);  // [SB][NC]
    const int nc = E.num_control_channels;

    for (int64_t n0 = (int64_t)blockIdx.x * ENC_STATES_PER_BLOCK; n0 < n;
         n0 += (int64_t)gridDim.x * ENC_STATES_PER_BLOCK) {
        const int nb = (int)((n - n0) < ENC_STATES_PER_BLOCK ? (n - n0) : ENC_STATES_PER_BLOCK);
        for (int i = threadIdx.x; i < nb * sw; i += ENC_THREADS) {
            const int s = i / sw, w = i - s * sw;
            cur[i] = states[(n0 + s) * sw + w];
            uint32_t p = 0;
            if (prev && E.num_prev_states > 0) {
                const int np = prev_count ? prev_count[n0 + s] : max_prev;
                for (int k = 0; k < np && k < max_prev; ++k)
                    p |= prev[((n0 + s) * max_prev + k) * sw + w];
            }
            por[i] = p;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nb * nc; i += ENC_THREADS) {
            const int s = i / nc, c = i - s * nc;
            ctrl[s * nc + c] = enc_control_value(E, cur + s * sw, c);
        }
        __syncthreads();
        // c_pad == 0: rows are [C][H][W] (NCHW, the reference layout). c_pad > 0: rows are
        // [H][W][c_pad] (channels-last with the channel count padded with zero planes), the layout the
        // tower's first convolution consumes without any conversion pass.
        const int row_elems = c_pad > 0 ? E.channel_size * c_pad : E.total;
        const int n_ch = E.total / E.channel_size;
        const int64_t flat0 = n0 * row_elems;
        const int64_t flat_n = (int64_t)nb * row_elems;
        // n0 is a multiple of 8 and rows are E.total elements: flat0 is a multiple of 8 elements
        for (int64_t v = threadIdx.x; v * 8 < flat_n; v += ENC_THREADS) {
            OutT vals[8];
            const int64_t e0 = v * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int64_t e = e0 + k;
                float x = 0.0f;
                if (e < flat_n) {
                    const int s = (int)(e / row_elems);
                    int le = (int)(e - (int64_t)s * row_elems);
                    bool live = true;
                    if (c_pad > 0) {
                        const int pos = le / c_pad, c = le - pos * c_pad;
                        live = c < n_ch;
                        le = c * E.channel_size + pos;
                    }
                    if (live)
                        x = enc_element(E, cur + s * sw, (prev && E.num_prev_states > 0) ? por + s * sw : nullptr,
                                        ctrl + s * nc, le);
                }
                vals[k] = enc_cast<OutT>(x);
            }
            OutT* dst = out + flat0 + e0;
            if (e0 + 8 <= flat_n) {
                if (sizeof(OutT) == 2) {
                    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<uint4*>(vals);
                } else {
                    reinterpret_cast<uint4*>(dst)[0] = reinterpret_cast<uint4*>(vals)[0];
                    reinterpret_cast<uint4*>(dst)[1] = reinterpret_cast<uint4*>(vals)[1];
                }
            } else {
                for (int k = 0; k < 8 && e0 + k < flat_n; ++k) dst[k] = vals[k];
            }
        }
        __syncthreads();
    }
}

// Fast path (any row size whose staging fits in shared memory): the per-element
// decisions (which base bit, current or previous state, which control channel) do not depend on the
// state, so a CTA resolves them ONCE into a descriptor table in shared memory and then streams states
// through it: per 8 output elements one descriptor read each, a bit test on the state words staged in
// shared memory, one 128-bit (bf16) or two 128-bit (f32) coalesced stores. No divisions per element.
#define ENCF_STATES 32
#define ENCF_THREADS 256
// descriptor: kind in bits 30-31 (0 zero, 1 bit of the state, 2 bit of the OR of previous states,
// 3 control channel), payload in bits 0-29 (base index / control channel)
template <typename OutT, bool STRADDLE /* rows are not a multiple of 8 elements: vectors may span two states */>
__global__ void __launch_bounds__(ENCF_THREADS)
encode_fast_kernel(const __grid_constant__ EncParams E, int64_t n, int sw, const uint32_t* __restrict__ states,
                   const uint32_t* __restrict__ prev, const uint8_t* __restrict__ prev_count, int max_prev,
                   OutT* __restrict__ out, int c_pad, int row_elems, int spb /* states per block iteration, multiple of 8 */) {
    extern __shared__ __align__(16) uint32_t fsm[];
    uint32_t* desc = fsm;                                       // [row_elems]
    uint32_t* cur = desc + row_elems;                           // [spb][sw]
    uint32_t* por = cur + spb * sw;                     // [spb][sw]
    uint32_t* ctrl = por + spb * sw;                    // [spb][nc] output bit patterns
    const int nc = E.num_control_channels;
    const int n_ch = E.total / E.channel_size;
    const bool has_prev = prev && E.num_prev_states > 0;
    for (int le0 = threadIdx.x; le0 < row_elems; le0 += ENCF_THREADS) {
        int le = le0;
        uint32_t d = 0u;
        bool live = true;
        if (c_pad > 0) {
            const int pos = le / c_pad, c = le - pos * c_pad;
            live = c < n_ch;
            le = c * E.channel_size + pos;
        }
        if (live) {
            if (le < E.board_total) {
                const int g = le / E.group_size;
                const int b = E.src_base[le - g * E.group_size];
                if (b >= 0 && g == 0) d = (1u << 30) | (uint32_t)b;
                else if (b >= 0 && g == 1 && has_prev) d = (2u << 30) | (uint32_t)b;
            } else {
                d = (3u << 30) | (uint32_t)((le - E.board_total) / E.channel_size);
            }
        }
        desc[le0] = d;
    }
    constexpr uint32_t ONE = sizeof(OutT) == 2 ? 0x3F80u : 0x3F800000u;
    for (int64_t n0 = (int64_t)blockIdx.x * spb; n0 < n; n0 += (int64_t)gridDim.x * spb) {
        const int nb = (int)((n - n0) < spb ? (n - n0) : spb);
        __syncthreads();
        for (int i = threadIdx.x; i < nb * sw; i += ENCF_THREADS) {
            const int s = i / sw, w = i - s * sw;
            cur[i] = states[(n0 + s) * sw + w];
            uint32_t p = 0;
            if (has_prev) {
                const int np = prev_count ? prev_count[n0 + s] : max_prev;
                for (int k = 0; k < np && k < max_prev; ++k) p |= prev[((n0 + s) * max_prev + k) * sw + w];
            }
            por[i] = p;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nb * nc; i += ENCF_THREADS) {
            const int s = i / nc, c = i - s * nc;
            const float v = enc_control_value(E, cur + s * sw, c);
            ctrl[i] = sizeof(OutT) == 2 ? (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v)) : __float_as_uint(v);
        }
        __syncthreads();
        // the block's rows are one contiguous run of nb * row_elems elements starting at a 16-byte boundary
        // (n0 is a multiple of 8): 8-element vectors over that run, a vector may straddle two states
        const int flat_n = nb * row_elems;
        const int items = flat_n >> 3;
        for (int w = threadIdx.x; w < items; w += ENCF_THREADS) {
            int s = (w * 8) / row_elems;
            int le = w * 8 - s * row_elems;
            uint32_t v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t* cs = cur + s * sw;
                const uint32_t* ps = por + s * sw;
                const uint32_t d = desc[STRADDLE ? le : le + k];
                const uint32_t kind = d >> 30, pl = d & 0x3FFFFFFFu;
                uint32_t x = 0u;
                if (kind == 1u) x = ((cs[pl >> 5] >> (pl & 31)) & 1u) ? ONE : 0u;
                else if (kind == 2u) x = ((ps[pl >> 5] >> (pl & 31)) & 1u) ? ONE : 0u;
                else if (kind == 3u) x = ctrl[s * nc + pl];
                v[k] = x;
                if (STRADDLE) { if (++le == row_elems) { le = 0; ++s; } }
            }
            OutT* dst = out + n0 * (int64_t)row_elems + (int64_t)w * 8;
            if (sizeof(OutT) == 2) {
                *reinterpret_cast<uint4*>(dst) = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16),
                                                            v[4] | (v[5] << 16), v[6] | (v[7] << 16));
            } else {
                reinterpret_cast<uint4*>(dst)[0] = make_uint4(v[0], v[1], v[2], v[3]);
                reinterpret_cast<uint4*>(dst)[1] = make_uint4(v[4], v[5], v[6], v[7]);
            }
        }
        // fewer than 8 elements can remain at the end of the batch's last block
        for (int e = (items << 3) + threadIdx.x; e < flat_n; e += ENCF_THREADS) {
            const int s = e / row_elems, le = e - s * row_elems;
            const uint32_t d = desc[le];
            const uint32_t kind = d >> 30, pl = d & 0x3FFFFFFFu;
            uint32_t x = 0u;
            if (kind == 1u) x = ((cur[s * sw + (pl >> 5)] >> (pl & 31)) & 1u) ? ONE : 0u;
            else if (kind == 2u) x = ((por[s * sw + (pl >> 5)] >> (pl & 31)) & 1u) ? ONE : 0u;
            else if (kind == 3u) x = ctrl[s * nc + pl];
            if (sizeof(OutT) == 2) reinterpret_cast<uint16_t*>(out)[n0 * (int64_t)row_elems + e] = (uint16_t)x;
            else reinterpret_cast<uint32_t*>(out)[n0 * (int64_t)row_elems + e] = x;
        }
    }
}

// ------------------------------------------------------------------------------------------------
extern "C" int gz_transformer_create(int channel_size, int channels_per_state, int num_control_channels,
                                     int num_prev_states, int num_rewards, const int* policy_sizes,
                                     int n_policies, gz_transformer** out) {
    if (!out || channel_size <= 0 || channels_per_state <= 0 || num_control_channels < 0 ||
        num_prev_states < 0 || n_policies < 0 || (n_policies > 0 && !policy_sizes)) {
        gz_set_error("bad transformer shape");
        return GZ_ERR_ARG;
    }
    gz_transformer* t = new gz_transformer();
    t->channel_size = channel_size;
    t->channels_per_state = channels_per_state;
    t->num_control_channels = num_control_channels;
    t->num_prev_states = num_prev_states;
    t->num_rewards = num_rewards;
    t->policy_sizes.assign(policy_sizes, policy_sizes + n_policies);
    *out = t;
    return GZ_OK;
}

extern "C" void gz_transformer_destroy(gz_transformer* t) {
    if (!t) return;
    if (t->d_arena) { cudaSetDevice(t->device); cudaFree(t->d_arena); }
    delete t;
}

extern "C" int gz_transformer_add_board_base(gz_transformer* t, int base_idx, int buf_offset) {
    if (!t || t->finalized) { gz_set_error("transformer is null or already finalised"); return GZ_ERR_ARG; }
    const int group = t->channels_per_state * t->channel_size;
    if (base_idx < 0 || buf_offset < 0 || buf_offset >= group) {
        gz_set_error("board base %d: offset %d outside one state's planes (%d)", base_idx, buf_offset, group);
        return GZ_ERR_ARG;
    }
    t->board_base.push_back(base_idx);
    t->board_off.push_back(buf_offset);
    return GZ_OK;
}

extern "C" int gz_transformer_add_control_base(gz_transformer* t, int base_idx, int channel_id, float value) {
    if (!t || t->finalized) { gz_set_error("transformer is null or already finalised"); return GZ_ERR_ARG; }
    if (base_idx < 0 || channel_id < 0 || channel_id >= t->num_control_channels) {
        gz_set_error("control base %d: channel %d out of range", base_idx, channel_id);
        return GZ_ERR_ARG;
    }
    t->ctrl_base.push_back(base_idx);
    t->ctrl_channel.push_back(channel_id);
    t->ctrl_value.push_back(value);
    return GZ_OK;
}

extern "C" int gz_transformer_total_size(const gz_transformer* t) {
    if (!t) return -1;
    return t->channel_size * (t->channels_per_state * (t->num_prev_states + 1) + t->num_control_channels);
}

extern "C" int gz_transformer_finalize(gz_transformer* t, int n_bases, int device) {
    if (!t) { gz_set_error("null transformer"); return GZ_ERR_ARG; }
    if (t->finalized) return GZ_OK;
    const int group = t->channels_per_state * t->channel_size;
    std::vector<int> src(group, -1);
    for (size_t i = 0; i < t->board_base.size(); ++i) {
        if (t->board_base[i] >= n_bases) { gz_set_error("board base index %d >= #bases %d", t->board_base[i], n_bases); return GZ_ERR_ARG; }
        if (src[t->board_off[i]] >= 0 && src[t->board_off[i]] != t->board_base[i]) {
            gz_set_error("two board bases (%d, %d) map to plane cell %d; unsupported", src[t->board_off[i]],
                         t->board_base[i], t->board_off[i]);
            return GZ_ERR_LIMIT;
        }
        src[t->board_off[i]] = t->board_base[i];
    }
    for (int b : t->ctrl_base)
        if (b >= n_bases) { gz_set_error("control base index %d >= #bases %d", b, n_bases); return GZ_ERR_ARG; }
    GZ_CUDA(cudaSetDevice(device));
    const size_t nctl = t->ctrl_base.size();
    const size_t bytes = (group + 2 * nctl + 8) * sizeof(int) + (nctl + 8) * sizeof(float);
    GZ_CUDA(cudaMalloc(&t->d_arena, bytes));
    int* d = (int*)t->d_arena;
    GZ_CUDA(cudaMemcpy(d, src.data(), group * sizeof(int), cudaMemcpyHostToDevice));
    int* d_cb = d + group;
    int* d_cc = d_cb + nctl;
    float* d_cv = (float*)(d_cc + nctl);
    if (nctl) {
        GZ_CUDA(cudaMemcpy(d_cb, t->ctrl_base.data(), nctl * sizeof(int), cudaMemcpyHostToDevice));
        GZ_CUDA(cudaMemcpy(d_cc, t->ctrl_channel.data(), nctl * sizeof(int), cudaMemcpyHostToDevice));
        GZ_CUDA(cudaMemcpy(d_cv, t->ctrl_value.data(), nctl * sizeof(float), cudaMemcpyHostToDevice));
    }
    EncParams& E = t->E;
    E.src_base = d;
    E.ctrl_base = d_cb;
    E.ctrl_channel = d_cc;
    E.ctrl_value = d_cv;
    E.n_control = (int)nctl;
    E.channel_size = t->channel_size;
    E.group_size = group;
    E.num_prev_states = t->num_prev_states;
    E.num_control_channels = t->num_control_channels;
    E.board_total = group * (t->num_prev_states + 1);
    E.total = gz_transformer_total_size(t);
    t->device = device;
    t->n_bases = n_bases;
    t->finalized = true;
    cudaDeviceProp prop;
    GZ_CUDA(cudaGetDeviceProperties(&prop, device));
    t->n_sms = prop.multiProcessorCount;
    return GZ_OK;
}

#include "gz_encoder_col.cuh"

// batches of at least this many states go to encode_col_kernel (gz_encoder_col.cuh); smaller ones — the PUCT step —
// keep encode_fast_kernel / encode_kernel
static int64_t g_col_min_batch = 16384;
extern "C" int gz_encode_set_column_kernel_min_batch(int64_t min_batch) {
    g_col_min_batch = min_batch < 0 ? ((int64_t)1 << 62) : min_batch;
    return GZ_OK;
}
// rows that are not a multiple of 8 elements through the column kernel too (units of states): -1 (default) where a
// tile fits the 48 KB every kernel may use (breakthroughSmall: measured 3.1x the small-batch kernels), 1 always
// (hexLG13's 3.7 KB states need the opt-in limit: bit-exact on a B200, not yet measured), 0 never
static int g_col_ragged = -1;
extern "C" int gz_encode_set_column_kernel_ragged(int on) {
    g_col_ragged = on < 0 ? -1 : (on ? 1 : 0);
    return GZ_OK;
}
static int g_col_tile = 0;          // states per tile, 0: the default (see gz_encode_layout)
extern "C" int gz_encode_set_column_kernel_tile(int tile_states) {
    g_col_tile = tile_states;
    return GZ_OK;
}

extern "C" int gz_encode_layout(gz_transformer* t, int64_t n, int state_words, const uint32_t* d_states,
                                const uint32_t* d_prev, const uint8_t* d_prev_count, int max_prev, void* d_out,
                                int out_dtype, int c_pad, void* stream);

extern "C" int gz_encode(gz_transformer* t, int64_t n, int state_words, const uint32_t* d_states,
                         const uint32_t* d_prev, const uint8_t* d_prev_count, int max_prev, void* d_out,
                         int out_dtype, void* stream) {
    return gz_encode_layout(t, n, state_words, d_states, d_prev, d_prev_count, max_prev, d_out, out_dtype, 0, stream);
}

extern "C" int gz_encode_layout(gz_transformer* t, int64_t n, int state_words, const uint32_t* d_states,
                                const uint32_t* d_prev, const uint8_t* d_prev_count, int max_prev, void* d_out,
                                int out_dtype, int c_pad, void* stream) {
    if (!t || !t->finalized) { gz_set_error("transformer is null or not finalised"); return GZ_ERR_ARG; }
    if (n < 0 || state_words * 32 < t->n_bases) { gz_set_error("bad batch / state_words"); return GZ_ERR_ARG; }
    if (n == 0) return GZ_OK;
    if (!d_states || !d_out) { gz_set_error("null device pointer"); return GZ_ERR_ARG; }
    if (d_prev && max_prev < 1) { gz_set_error("max_prev must be >= 1 with previous states"); return GZ_ERR_ARG; }
    if (c_pad != 0 && (c_pad < t->E.total / t->E.channel_size || (c_pad & 7))) { gz_set_error("c_pad must be a multiple of 8 and >= the channel count"); return GZ_ERR_ARG; }
    if ((((uintptr_t)d_out) & 15) != 0) { gz_set_error("plane buffer must be 16-byte aligned"); return GZ_ERR_ARG; }
    GZ_CUDA(cudaSetDevice(t->device));
    const int row_elems_c = c_pad > 0 ? t->E.channel_size * c_pad : t->E.total;
    if (n >= g_col_min_batch && (out_dtype == GZ_DTYPE_BF16 || out_dtype == GZ_DTYPE_F32) && row_elems_c >= 8 &&
        ((row_elems_c & 7) == 0 || g_col_ragged != 0)) {
        const int row_elems = row_elems_c;
        // rows that are not a multiple of 8 elements go in units of U states (gz_encoder_col.cuh)
        int U = 1;
        while ((U * row_elems) & 7) U <<= 1;
        const int BW = 2 * state_words + t->num_control_channels + 1;
        // the tile (measured, profiles/r02_encoder_column_kernel.jsonl): what matters is the number of 8-element
        // chunks a thread stores per tile — about 40 for bf16 rows, about 12 for f32 rows
        const bool half = out_dtype == GZ_DTYPE_BF16;
        int S = g_col_tile;
        if (g_col_tile < 8) {
            const int CPR = (U * row_elems) >> 3;
            const int NT = CPR >= ENCC_THREADS ? ENCC_THREADS : (ENCC_THREADS / CPR) * CPR;
            const double want = (half ? 40.0 : 12.0) * NT / CPR * U;       // states
            S = 8;
            while (S < 256 && (double)S * 1.4142 < want) S <<= 1;          // nearest power of two
            const int lo = 32, hi = half ? 256 : 64;
            S = S < lo ? lo : (S > hi ? hi : S);
        }
        if (g_col_tile < 8)                      // mid-size batches: enough tiles for every SM's resident CTAs
            while (S > 8 && (n + S - 1) / S < (int64_t)t->n_sms * 4) S >>= 1;
        auto need = [&](int k) { return ((size_t)U * row_elems + (size_t)k * BW) * 4; };
        const size_t cap = (U > 1 && g_col_ragged == 1) ? 160 * 1024 : 48 * 1024;     // (large states — hexLG13 — need the opt-in limit)
        while (S > 8 && need(S) > cap) S >>= 1;
        if (S >= U && (S % U) == 0 && need(S) <= cap) {
            const int64_t blocks = (n + S - 1) / S;
            const int grid = (int)std::min<int64_t>(blocks, (int64_t)t->n_sms * 8);
            if (out_dtype == GZ_DTYPE_BF16) {
                if (need(S) > 48 * 1024)
                    GZ_CUDA(cudaFuncSetAttribute(encode_col_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need(S)));
                encode_col_kernel<__nv_bfloat16><<<grid, ENCC_THREADS, need(S), (cudaStream_t)stream>>>(
                    t->E, n, state_words, d_states, d_prev, d_prev_count, max_prev, (__nv_bfloat16*)d_out, c_pad, row_elems, S, U);
            } else {
                if (need(S) > 48 * 1024)
                    GZ_CUDA(cudaFuncSetAttribute(encode_col_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need(S)));
                encode_col_kernel<float><<<grid, ENCC_THREADS, need(S), (cudaStream_t)stream>>>(
                    t->E, n, state_words, d_states, d_prev, d_prev_count, max_prev, (float*)d_out, c_pad, row_elems, S, U);
            }
            GZ_CUDA(cudaGetLastError());
            return GZ_OK;
        }
    }
    {
        const int row_elems = c_pad > 0 ? t->E.channel_size * c_pad : t->E.total;
        int spb = ENCF_STATES;                 // (states of hexLG13's size do not fit: generic kernel)
        auto need = [&](int k) { return ((size_t)row_elems + (size_t)k * (2 * state_words + t->num_control_channels)) * 4 + 16; };
        const size_t fsmem = need(spb);
        const bool straddle = (row_elems & 7) != 0;
        if (fsmem <= 200 * 1024 && (out_dtype == GZ_DTYPE_BF16 || out_dtype == GZ_DTYPE_F32)) {
            const int64_t blocks = (n + spb - 1) / spb;
            const int grid = (int)std::min<int64_t>(blocks, (int64_t)t->n_sms * 8);
#define GZ_ENCF_LAUNCH(T, ST)                                                                                   \
    do {                                                                                                        \
        if (fsmem > 48 * 1024)                                                                                  \
            GZ_CUDA(cudaFuncSetAttribute(encode_fast_kernel<T, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem)); \
        encode_fast_kernel<T, ST><<<grid, ENCF_THREADS, fsmem, (cudaStream_t)stream>>>(                            \
            t->E, n, state_words, d_states, d_prev, d_prev_count, max_prev, (T*)d_out, c_pad, row_elems, spb);  \
    } while (0)
            if (out_dtype == GZ_DTYPE_BF16) {
                if (straddle) GZ_ENCF_LAUNCH(__nv_bfloat16, true); else GZ_ENCF_LAUNCH(__nv_bfloat16, false);
            } else {
                if (straddle) GZ_ENCF_LAUNCH(float, true); else GZ_ENCF_LAUNCH(float, false);
            }
#undef GZ_ENCF_LAUNCH
            GZ_CUDA(cudaGetLastError());
            return GZ_OK;
        }
    }
    const size_t smem = (size_t)ENC_STATES_PER_BLOCK * (2 * state_words * 4 + t->num_control_channels * 4) + 16;
    const int64_t blocks = (n + ENC_STATES_PER_BLOCK - 1) / ENC_STATES_PER_BLOCK;
    const int grid = (int)std::min<int64_t>(blocks, (int64_t)t->n_sms * 8);
    if (out_dtype == GZ_DTYPE_BF16) {
        if (smem > 48 * 1024)
            GZ_CUDA(cudaFuncSetAttribute(encode_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        encode_kernel<__nv_bfloat16><<<grid, ENC_THREADS, smem, (cudaStream_t)stream>>>(
            t->E, n, state_words, d_states, d_prev, d_prev_count, max_prev, (__nv_bfloat16*)d_out, c_pad);
    } else if (out_dtype == GZ_DTYPE_F32) {
        if (smem > 48 * 1024)
            GZ_CUDA(cudaFuncSetAttribute(encode_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        encode_kernel<float><<<grid, ENC_THREADS, smem, (cudaStream_t)stream>>>(
            t->E, n, state_words, d_states, d_prev, d_prev_count, max_prev, (float*)d_out, c_pad);
    } else {
        gz_set_error("bad out_dtype");
        return GZ_ERR_ARG;
    }
    GZ_CUDA(cudaGetLastError());
    return GZ_OK;
}

// ------------------------------------------------------------------------------------------------
// Board symmetries of packed states (src/ggpzero/util/symmetry.py translate_basestate, batched):
// out bit j of every state = in bit src[j] — one bit-permutation shared by the whole batch. A CTA
// stages 32 states in shared memory; a thread assembles one output word from 32 gathered bits.
#define PERM_STATES 32
__global__ void __launch_bounds__(256)
permute_state_bits_kernel(int64_t n, int sw, int n_bases, const uint32_t* __restrict__ states,
                          const int* __restrict__ src, uint32_t* __restrict__ out) {
    extern __shared__ __align__(16) uint32_t psm[];
    for (int64_t n0 = (int64_t)blockIdx.x * PERM_STATES; n0 < n; n0 += (int64_t)gridDim.x * PERM_STATES) {
        const int nb = (int)((n - n0) < PERM_STATES ? (n - n0) : PERM_STATES);
        __syncthreads();
        for (int i = threadIdx.x; i < nb * sw; i += blockDim.x) psm[i] = states[n0 * sw + i];
        __syncthreads();
        for (int i = threadIdx.x; i < nb * sw; i += blockDim.x) {
            const int s = i / sw, w = i - s * sw;
            const uint32_t* row = psm + s * sw;
            uint32_t v = 0u;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
                const int j = w * 32 + k;
                if (j < n_bases) {
                    const int b = __ldg(src + j);
                    v |= ((row[b >> 5] >> (b & 31)) & 1u) << k;
                }
            }
            out[n0 * sw + i] = v;
        }
    }
}

extern "C" int gz_permute_state_bits(int64_t n, int state_words, int n_bases, const uint32_t* d_states,
                                     const int* d_src, uint32_t* d_out, void* stream) {
    if (n < 0 || state_words <= 0 || n_bases <= 0 || n_bases > state_words * 32) { gz_set_error("bad shape"); return GZ_ERR_ARG; }
    if (n == 0) return GZ_OK;
    if (!d_states || !d_src || !d_out || d_states == d_out) { gz_set_error("null or aliased device pointer"); return GZ_ERR_ARG; }
    const size_t smem = (size_t)PERM_STATES * state_words * 4;
    if (smem > 200 * 1024) { gz_set_error("state rows too large for the staging block"); return GZ_ERR_LIMIT; }
    if (smem > 48 * 1024)
        GZ_CUDA(cudaFuncSetAttribute(permute_state_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t blocks = (n + PERM_STATES - 1) / PERM_STATES;
    permute_state_bits_kernel<<<(int)std::min<int64_t>(blocks, 148 * 8), 256, smem, (cudaStream_t)stream>>>(
        n, state_words, n_bases, d_states, d_src, d_out);
    GZ_CUDA(cudaGetLastError());
    return GZ_OK;
}

