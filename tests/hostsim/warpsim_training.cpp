// warpsim_training.cpp — CPU validation build of rec_unpack_kernel (csrc/gz_training_kernel.cuh). TEST
// INFRASTRUCTURE: the kernel source is compiled unchanged for the host; a warp is 32 std::threads that meet at a
// barrier inside every warp-level primitive (__ballot_sync / __shfl_xor_sync / __any_sync), so divergent lanes,
// lane-strided loops and the ballots behave as on the device (a lane that skipped a primitive would deadlock the
// warp — the driver in tests/test_training_batch.py runs this under a timeout). The kernel uses no CTA-level
// synchronisation, so the warps of the grid are simply run one after the other.
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

struct Dim3 { int x, y, z; };
static thread_local Dim3 threadIdx, blockIdx;
static Dim3 gridDim, blockDim;

struct WarpCtx {
    std::barrier<> bar{32};
    uint32_t u[32];
    double d[32];
};
static thread_local WarpCtx* W;
static thread_local int LANE;

static inline uint32_t __ballot_sync(unsigned, bool p) {
    W->u[LANE] = p ? 1u : 0u;
    W->bar.arrive_and_wait();
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i) r |= W->u[i] << i;
    W->bar.arrive_and_wait();
    return r;
}
static inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0u; }
static inline double __shfl_xor_sync(unsigned, double v, int lane_mask) {
    W->d[LANE] = v;
    W->bar.arrive_and_wait();
    const double r = W->d[LANE ^ lane_mask];
    W->bar.arrive_and_wait();
    return r;
}
static inline float __int_as_float(int x) { float f; memcpy(&f, &x, 4); return f; }
static inline int atomicAdd(int32_t* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

#include "../../galvanise_zero_b200/csrc/gz_training_kernel.cuh"

extern "C" int ws_records_to_training(int64_t n, int64_t n_records, int record_words, int state_words, int n_bases,
                                      int n_roles, const int* widths, int n_values, const int32_t* records,
                                      const int32_t* index, const int32_t* sym, int n_sym, const int32_t* base_src,
                                      const int32_t* const* action_dest, float* const* policy, float* value,
                                      uint32_t* states, uint32_t* prev, uint8_t* prev_count, int32_t* bad, int grid) {
    TrnArgs A{};
    A.n = n; A.n_records = n_records; A.record_words = record_words; A.state_words = state_words;
    A.n_bases = n_bases; A.n_roles = n_roles; A.n_values = n_values; A.n_sym = sym ? n_sym : 0;
    A.records = records; A.index = index; A.sym = sym; A.base_src = base_src;
    for (int r = 0; r < n_roles; ++r) {
        A.policy[r] = policy[r];
        A.width[r] = widths[r];
        A.action_dest[r] = sym ? action_dest[r] : nullptr;
        memset(policy[r], 0, (size_t)n * widths[r] * sizeof(float));     // (cudaMemsetAsync in gz_records_to_training)
    }
    A.value = value; A.states = states; A.prev = prev; A.prev_count = prev_count; A.bad = bad;
    if (bad) *bad = 0;
    gridDim = {grid, 1, 1};
    blockDim = {TRN_WARPS * 32, 1, 1};
    for (int b = 0; b < grid; ++b)
        for (int w = 0; w < TRN_WARPS; ++w) {
            WarpCtx ctx;
            std::vector<std::thread> lanes;
            for (int l = 0; l < 32; ++l)
                lanes.emplace_back([&, b, w, l] {
                    W = &ctx; LANE = l;
                    blockIdx = {b, 0, 0};
                    threadIdx = {w * 32 + l, 0, 0};
                    rec_unpack_kernel(A);
                });
            for (auto& t : lanes) t.join();
        }
    return 0;
}
