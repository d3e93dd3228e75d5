// ctasim_encoder.cpp — CPU validation build of encode_col_kernel (csrc/gz_encoder_col.cuh). TEST INFRASTRUCTURE:
// the kernel source is compiled unchanged for the host; a CTA is ENCC_THREADS std::threads that meet at a
// std::barrier wherever the kernel calls __syncthreads(), the CTAs of the grid run one after the other. Driven by
// tests/test_encoder.py against the C oracle encoder. bf16 rounding is CUDA's own host implementation
// (cuda_bf16.h compiles for the host).
#include <cuda_bf16.h>
#include <vector_functions.h>
#include <vector_types.h>

#include <barrier>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#undef __global__
#undef __device__
#undef __forceinline__
#undef __launch_bounds__
#undef __grid_constant__
#undef __restrict__
#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __restrict__

struct SimDim3 { int x, y, z; };
static thread_local SimDim3 sim_threadIdx, sim_blockIdx;
static SimDim3 sim_gridDim;
#define threadIdx sim_threadIdx
#define blockIdx sim_blockIdx
#define gridDim sim_gridDim

static std::barrier<>* g_cta_barrier;
static uint32_t* g_cta_smem;
static inline void sim_syncthreads() { g_cta_barrier->arrive_and_wait(); }
#define __syncthreads sim_syncthreads
static inline uint32_t sim_float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
#define __float_as_uint sim_float_as_uint
#define ENCC_DYNAMIC_SMEM(name) uint32_t* name = g_cta_smem
#define ENCC_HOST_BUILD 1

#include "../../galvanise_zero_b200/csrc/gz_encoder_col.cuh"

// out_dtype: 0 float32, 1 bf16 (GZ_DTYPE_*). Tables as gz_transformer_finalize builds them.
extern "C" int cs_encode_col(const int* src_base, int group_size, const int* ctrl_base, const int* ctrl_channel,
                             const float* ctrl_value, int n_control, int channel_size, int num_prev_states,
                             int num_control_channels, int64_t n, int sw, const uint32_t* states,
                             const uint32_t* prev, const uint8_t* prev_count, int max_prev, void* out, int out_dtype,
                             int c_pad, int S, int grid) {
    EncParams E{};
    E.src_base = src_base; E.ctrl_base = ctrl_base; E.ctrl_channel = ctrl_channel; E.ctrl_value = ctrl_value;
    E.n_control = n_control; E.channel_size = channel_size; E.group_size = group_size;
    E.num_prev_states = num_prev_states; E.num_control_channels = num_control_channels;
    E.board_total = group_size * (num_prev_states + 1);
    E.total = E.board_total + num_control_channels * channel_size;
    const int row_elems = c_pad > 0 ? channel_size * c_pad : E.total;
    int U = 1;
    while ((U * row_elems) & 7) U <<= 1;
    if (S % U) return -1;
    const int BW = 2 * sw + num_control_channels + 1;
    std::vector<uint32_t> smem((size_t)U * row_elems + (size_t)S * BW);
    sim_gridDim = {grid, 1, 1};
    for (int b = 0; b < grid; ++b) {
        std::barrier<> bar(ENCC_THREADS);
        g_cta_barrier = &bar;
        g_cta_smem = smem.data();
        std::vector<std::thread> th;
        for (int t = 0; t < ENCC_THREADS; ++t)
            th.emplace_back([&, b, t] {
                sim_blockIdx = {b, 0, 0};
                sim_threadIdx = {t, 0, 0};
                if (out_dtype == 1)
                    encode_col_kernel<__nv_bfloat16>(E, n, sw, states, prev, prev_count, max_prev, (__nv_bfloat16*)out,
                                                     c_pad, row_elems, S, U);
                else
                    encode_col_kernel<float>(E, n, sw, states, prev, prev_count, max_prev, (float*)out, c_pad,
                                             row_elems, S, U);
            });
        for (auto& x : th) x.join();
    }
    return 0;
}
