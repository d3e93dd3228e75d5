// gz_spec.cuh — game-specialised, THREAD-sliced propnet kernels for large batches (sm_100a).
//
// pn_kernel (gz_propnet.cu) interprets gate tables with one warp per 32-game tile: lane = gate, one
// 32-bit word = 32 games. It is bound by instruction issue (profiles/r01_propnet_playout_ncu_full.md:
// 4.7 K warp instructions per tile, most of them literal decoding, shared-memory traffic and shuffle
// transposes). Here one THREAD owns a tile: the same bit-sliced words live in registers, the gate
// network of the game is straight-line code emitted by tools/gen_spec_kernel.py from the compiled
// propnet (csrc/gz_spec_<game>.cu), 32x32 transposes run inside the thread, and a warp moves its 32
// tiles (1024 games, contiguous in HBM) with coalesced 128-bit accesses through a padded
// shared-memory block. A warp instruction now advances 32 tiles instead of one.
//
// Same contract and the same results as pn_kernel (tests/test_gpu_propnet.py runs both against the C
// oracle); used by gz_sm_* for batches of at least gz_propnet_set_spec_min_batch() states.
#ifndef GZ_SPEC_CUH
#define GZ_SPEC_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>
#include <utility>

#include "../../include/gz_b200.h"
#include "../../include/gz_propnet_format.h"

enum { SPEC_EVAL = 0, SPEC_TRANSITION = 1, SPEC_PLAYOUT = 2 };

struct SpecArgs {
    int64_t n, slot0;
    const uint32_t* states_in;
    const int16_t* moves_in;
    uint32_t* states_out;
    uint32_t* legal;
    uint8_t* terminal;
    uint8_t* goals;
    int16_t* moves_out;
    unsigned long long* transitions;
    uint64_t seed, step;
    int goals_mode;
};

// In-register transpose of a 32x32 bit matrix, LSB-first: bit c of a[r] <-> bit r of a[c].
// Block-swap butterfly: the 16- and 8-bit stages are byte permutes (2 PRMT per pair), the 4/2/1-bit stages
// a shift + LOP3 per output word.
__device__ __forceinline__ void spec_transpose32(uint32_t (&a)[32]) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const uint32_t x = a[k], y = a[k + 16];
        a[k] = __byte_perm(x, y, 0x5410);
        a[k + 16] = __byte_perm(x, y, 0x7632);
    }
#pragma unroll
    for (int k = 0; k < 32; k = (k + 8 + 1) & ~8) {
        const uint32_t x = a[k], y = a[k + 8];
        a[k] = __byte_perm(x, y, 0x6240);
        a[k + 8] = __byte_perm(x, y, 0x7351);
    }
#pragma unroll
    for (int j = 4; j != 0; j >>= 1) {
        const uint32_t m = j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
#pragma unroll
        for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
            const uint32_t x = a[k], y = a[k + j];
            a[k] = (x & m) | ((y << j) & ~m);
            a[k + j] = ((x >> j) & m) | (y & ~m);
        }
    }
}

// index of the (k+1)-th set bit of x (k < popc(x)), branch-free binary search
__device__ __forceinline__ int spec_kth_bit(uint32_t x, uint32_t k) {
    int pos = 0;
    uint32_t c;
    c = __popc(x & 0xFFFFu);           if (k >= c) { k -= c; pos += 16; x >>= 16; }
    c = __popc(x & 0xFFu);             if (k >= c) { k -= c; pos += 8; x >>= 8; }
    c = __popc(x & 0xFu);              if (k >= c) { k -= c; pos += 4; x >>= 4; }
    c = __popc(x & 0x3u);              if (k >= c) { k -= c; pos += 2; x >>= 2; }
    c = x & 1u;                        if (k >= c) { pos += 1; }
    return pos;
}

// A warp's shared-memory block holds 32 tile blocks of BW words (BW % 32 == 0), each padded by 4 words:
// the coalesced fill / drain (consecutive lanes, consecutive 16-byte chunks) and every lane's 128-bit
// accesses to its own tile block are both bank-conflict free. The fill is cp.async (global -> shared
// without a register round trip): all of a warp's 16-byte chunks are in flight at once.
template <int BW>
__device__ __forceinline__ void spec_fill_async(uint32_t* buf, const uint32_t* __restrict__ src, int64_t rows_avail,
                                                int row_words, int lane) {
    const int64_t words_avail = rows_avail * row_words;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(buf);
#pragma unroll 8
    for (int q = lane; q < 32 * BW / 4; q += 32) {
        const int w = q * 4;
        const int t = w / BW, o = w - t * BW;
        const uint32_t dst = sbase + (uint32_t)(t * (BW + 4) + o) * 4u;
        const int bytes = (int64_t)w < words_avail ? 16 : 0;          // rows beyond the batch read as zero
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src + w), "r"(bytes) : "memory");
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void spec_fill_wait() {
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncwarp();
}

template <int BW>
__device__ __forceinline__ void spec_drain(const uint32_t* buf, uint32_t* __restrict__ dst, int64_t rows_avail,
                                           int row_words, int lane) {
    __syncwarp();
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    const int64_t words_avail = rows_avail * row_words;
#pragma unroll 8
    for (int q = lane; q < 32 * BW / 4; q += 32) {
        const int w = q * 4;
        const int t = w / BW, o = w - t * BW;
        if ((int64_t)w < words_avail) d4[q] = *reinterpret_cast<const uint4*>(buf + t * (BW + 4) + o);
    }
    __syncwarp();
}

// compile-time loop
template <int I, int N, typename F>
__device__ __forceinline__ void spec_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        spec_for<I + 1, N>(f);
    }
}

#define SPEC_WARPS 4

// G: the generated game description (csrc/gz_spec_<game>.cu), see tools/gen_spec_kernel.py.
//
// Code size matters here: the gate networks are straight-line by nature (a few thousand LOP3s that
// every warp runs once per batch), so everything repetitive — the per-game move chooser, the word-by-word
// transposes back to packed rows — is a real loop whose body stays in the instruction cache, and the
// four warps of a CTA enter the straight-line sections together (__syncthreads) so that one fetch of a
// line feeds all of them.
template <typename G, int MODE>
__global__ void __launch_bounds__(SPEC_WARPS * 32, 1) spec_kernel(const SpecArgs A) {
    extern __shared__ __align__(16) uint32_t spec_smem[];
    constexpr int SW = G::SW, LW = G::LW, R = G::R, NBW = G::NBW;
    constexpr int SBW = 32 * SW, LBW = 32 * LW;                  // words per tile block
    constexpr int BUFW = 32 * ((LBW > SBW ? LBW : SBW) + 4);     // words per warp buffer
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* buf = spec_smem + (size_t)warp * BUFW;
    uint32_t* mine_s = buf + lane * (SBW + 4);                   // this lane's tile block (state rows)
    uint32_t* mine_l = buf + lane * (LBW + 4);                   // ... (legal rows)

    const int64_t n = A.n;
    const int64_t n_batches = (n + 1023) >> 10;                  // a warp's unit: 32 tiles = 1024 games
    const int64_t per_round = (int64_t)gridDim.x * SPEC_WARPS;
    const int64_t rounds = (n_batches + per_round - 1) / per_round;
    for (int64_t round = 0; round < rounds; ++round) {
        const int64_t batch = round * per_round + (int64_t)blockIdx.x * SPEC_WARPS + warp;
        const bool active = batch < n_batches;                   // idle warps still meet the barriers
        const int64_t w0 = batch << 10;                          // first game of the warp
        const int64_t g0 = w0 + ((int64_t)lane << 5);            // first game of this lane's tile
        const int64_t warp_rows = !active ? 0 : (n - w0 < 1024 ? n - w0 : 1024);
        int rows = !active ? 0 : (int)(n - g0 < 32 ? n - g0 : 32);
        if (rows < 0) rows = 0;
        const uint32_t valid_mask = rows >= 32 ? 0xffffffffu : ((1u << rows) - 1u);

        uint32_t reset_mask = 0u, pass_mask = 0u;
        uint32_t mv[R][G::MVBITS];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int k = 0; k < G::MVBITS; ++k) mv[r][k] = 0u;

        if (MODE == SPEC_PLAYOUT && active) {
            // ---- one uniformly random legal action per role and game, from the stored legal rows
            spec_fill_async<LBW>(buf, A.legal + w0 * LW, warp_rows, LW, lane);
            if (rows == 32) {
                const uint4* t4 = reinterpret_cast<const uint4*>(A.terminal + g0);
                const uint4 ta = t4[0], tb2 = t4[1];
                const uint32_t tw8[8] = {ta.x, ta.y, ta.z, ta.w, tb2.x, tb2.y, tb2.z, tb2.w};
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    // one flag byte per game -> 4 bits
                    const uint32_t nz = (tw8[q] | (tw8[q] >> 4) | (tw8[q] >> 2) | (tw8[q] >> 1) |
                                         (tw8[q] >> 7) | (tw8[q] >> 6) | (tw8[q] >> 5) | (tw8[q] >> 3)) & 0x01010101u;
                    reset_mask |= (((nz * 0x10204080u) >> 28) & 0xFu) << (4 * q);
                }
            } else {
                for (int i = 0; i < rows; ++i) reset_mask |= (A.terminal[g0 + i] ? 1u : 0u) << i;
            }
            spec_fill_wait();
#pragma unroll 2
            for (int i = 0; i < 32; ++i) {
                const bool term = (reset_mask >> i) & 1u;
                const bool valid = i < rows;
                uint32_t lw[LW];
#pragma unroll
                for (int q = 0; q < LW / 4; ++q) {
                    const uint4 v = *reinterpret_cast<const uint4*>(mine_l + i * LW + q * 4);
                    lw[q * 4] = v.x; lw[q * 4 + 1] = v.y; lw[q * 4 + 2] = v.z; lw[q * 4 + 3] = v.w;
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int wlo = G::lwo(r), whi = G::lwo(r + 1);
                    uint32_t cnt = 0;
#pragma unroll
                    for (int j = wlo; j < whi; ++j) cnt += __popc(lw[j]);
                    uint32_t pick = gz_pick(gz_rand32(A.seed, (uint64_t)(A.slot0 + g0 + i), A.step, (uint32_t)r), cnt);
                    // the word holding the pick-th legal action, then its position inside the word
                    uint32_t xsel = 0u;
                    int base = 0;
                    bool found = false;
#pragma unroll
                    for (int j = wlo; j < whi; ++j) {
                        const uint32_t c = __popc(lw[j]);
                        const bool here = !found && pick < c;
                        if (here) { xsel = lw[j]; base = (j - wlo) * 32; }
                        found = found || here;
                        if (!found) pick -= c;
                    }
                    int idx = found ? base + spec_kth_bit(xsel, pick) : 0;
                    if (!valid || term) idx = 0;
                    if (valid && A.moves_out) A.moves_out[(g0 + i) * R + r] = (int16_t)(term ? -1 : idx);
#pragma unroll
                    for (int k = 0; k < G::MVBITS; ++k) mv[r][k] |= ((uint32_t)(idx >> k) & 1u) << i;
                }
            }
            __syncwarp();
        } else if (MODE == SPEC_TRANSITION && active) {
#pragma unroll 1
            for (int i = 0; i < rows; ++i) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    int idx = (int)A.moves_in[(g0 + i) * R + r];
                    if (r == 0 && idx < 0) pass_mask |= 1u << i;
                    if (idx < 0) idx = 0;
#pragma unroll
                    for (int k = 0; k < G::MVBITS; ++k) mv[r][k] |= ((uint32_t)(idx >> k) & 1u) << i;
                }
            }
        }

        // ---- packed state rows -> bit-sliced bases in registers
        uint32_t b[NBW * 32];
        if (active) spec_fill_async<SBW>(buf, A.states_in + w0 * SW, warp_rows, SW, lane);
        spec_fill_wait();
        __syncthreads();
        spec_for<0, NBW>([&](auto J) {
            constexpr int j = decltype(J)::value;
            uint32_t a[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = mine_s[i * SW + j];
            spec_transpose32(a);
#pragma unroll
            for (int i = 0; i < 32; ++i) b[j * 32 + i] = a[i];
        });
        __syncwarp();

        uint32_t sl[(G::LWU > NBW ? G::LWU : NBW) * 32];        // bit-sliced outputs, indexed in loops (local memory)

        if (MODE != SPEC_EVAL) {
            // ---- next state: straight-line gate network per state word ...
            typename G::Decode dec;
            G::decode_moves(mv, dec);
            __syncthreads();
            spec_for<0, NBW>([&](auto J) {
                constexpr int j = decltype(J)::value;
                uint32_t o[32];
                G::template next_word<j>(b, dec, o);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    uint32_t t = o[i];
                    if (MODE == SPEC_PLAYOUT) {
                        t = G::init_bit(j * 32 + i) ? (t | reset_mask) : (t & ~reset_mask);
                    } else {
                        t = (t & ~pass_mask) | (b[j * 32 + i] & pass_mask);
                    }
                    t &= valid_mask;
                    if (j * 32 + i >= G::NB) t = 0u;
                    sl[j * 32 + i] = t;
                }
            });
            // ... the new bases replace the old ones, and go back to packed rows word by word
#pragma unroll
            for (int i = 0; i < NBW * 32; ++i) b[i] = sl[i];
#pragma unroll 1
            for (int j = 0; j < NBW; ++j) {
                uint32_t o[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = sl[j * 32 + i];
                spec_transpose32(o);
#pragma unroll
                for (int i = 0; i < 32; ++i) mine_s[i * SW + j] = o[i];
            }
#pragma unroll
            for (int j = NBW; j < SW; ++j)
#pragma unroll
                for (int i = 0; i < 32; ++i) mine_s[i * SW + j] = 0u;
            if (active) spec_drain<SBW>(buf, A.states_out + w0 * SW, warp_rows, SW, lane);
        }

        // ---- views of the (new) state: legal masks, terminal, goals
        __syncthreads();
        spec_for<0, G::LWU>([&](auto J) {
            constexpr int j = decltype(J)::value;
            uint32_t o[32];
            G::template legal_word<j>(b, o);
#pragma unroll
            for (int i = 0; i < 32; ++i) sl[j * 32 + i] = o[i] & valid_mask;
        });
        const uint32_t tw = G::terminal(b) & valid_mask;
        const bool need_goals = (A.goals_mode == GZ_GOALS_ALWAYS) || (tw != 0u);
        uint32_t gw[G::NGOAL > 0 ? G::NGOAL : 1];
#pragma unroll
        for (int k = 0; k < (G::NGOAL > 0 ? G::NGOAL : 1); ++k) gw[k] = 0u;
        if (need_goals) G::goal_words(b, gw);
#pragma unroll 1
        for (int j = 0; j < G::LWU; ++j) {
            uint32_t o[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = sl[j * 32 + i];
            spec_transpose32(o);
#pragma unroll
            for (int i = 0; i < 32; ++i) mine_l[i * LW + j] = o[i];
        }
#pragma unroll
        for (int j = G::LWU; j < LW; ++j)
#pragma unroll
            for (int i = 0; i < 32; ++i) mine_l[i * LW + j] = 0u;
        if (active) spec_drain<LBW>(buf, A.legal + w0 * LW, warp_rows, LW, lane);

        if (rows == 32) {
            // 32 terminal bytes and 32 x R goal bytes per tile, written as 128-bit words
            uint32_t tb[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) tb[q] = (((tw >> (4 * q)) & 0xFu) * 0x00204081u) & 0x01010101u;
            uint4* t4 = reinterpret_cast<uint4*>(A.terminal + g0);
            t4[0] = make_uint4(tb[0], tb[1], tb[2], tb[3]);
            t4[1] = make_uint4(tb[4], tb[5], tb[6], tb[7]);
            uint32_t* gout = reinterpret_cast<uint32_t*>(A.goals + g0 * R);
#pragma unroll 1
            for (int q = 0; q < 32 * R / 4; ++q) {          // 4 goal bytes per word
                uint32_t wv = 0u;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int flat = q * 4 + e, i = flat / R, r = flat - i * R;
                    uint32_t gv = 255u;
#pragma unroll
                    for (int rr = 0; rr < R; ++rr)
#pragma unroll
                        for (int k = G::goal_off(rr); k < G::goal_off(rr + 1); ++k)
                            if (rr == r && ((gw[k] >> i) & 1u)) gv = G::goal_val(k);
                    if (!need_goals || (A.goals_mode == GZ_GOALS_TERMINAL_ONLY && !((tw >> i) & 1u))) gv = 255u;
                    wv |= gv << (8 * e);
                }
                gout[q] = wv;
            }
        } else {
#pragma unroll 1
            for (int i = 0; i < rows; ++i) {
                const bool is_term = (tw >> i) & 1u;
                A.terminal[g0 + i] = is_term ? 1 : 0;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    uint32_t gv = 255u;
#pragma unroll
                    for (int k = G::goal_off(r); k < G::goal_off(r + 1); ++k)
                        if ((gw[k] >> i) & 1u) gv = G::goal_val(k);
                    if (!need_goals || (A.goals_mode == GZ_GOALS_TERMINAL_ONLY && !is_term)) gv = 255u;
                    A.goals[(g0 + i) * R + r] = (uint8_t)gv;
                }
            }
        }
        if (MODE == SPEC_PLAYOUT && A.transitions) {
            int cnt = __popc(valid_mask & ~reset_mask);
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
            if (lane == 0 && cnt) atomicAdd(A.transitions, (unsigned long long)cnt);
        }
        __syncwarp();
    }
}

// host side: a specialised game registers itself by the hash of its device program
typedef int (*gz_spec_launch_fn)(int mode, const SpecArgs& a, int grid, cudaStream_t s);
struct GzSpecEntry {
    uint64_t hash;
    const char* game;
    gz_spec_launch_fn launch;
    GzSpecEntry* next;
};
void gz_spec_register(GzSpecEntry* e);
const GzSpecEntry* gz_spec_find(uint64_t hash);

template <typename G>
int gz_spec_launch_impl(int mode, const SpecArgs& a, int grid, cudaStream_t s) {
    constexpr int SBW = 32 * G::SW, LBW = 32 * G::LW;
    constexpr size_t smem = (size_t)SPEC_WARPS * 32 * ((LBW > SBW ? LBW : SBW) + 4) * 4;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(spec_kernel<G, SPEC_EVAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(spec_kernel<G, SPEC_TRANSITION>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(spec_kernel<G, SPEC_PLAYOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    if (mode == SPEC_EVAL) spec_kernel<G, SPEC_EVAL><<<grid, SPEC_WARPS * 32, smem, s>>>(a);
    else if (mode == SPEC_TRANSITION) spec_kernel<G, SPEC_TRANSITION><<<grid, SPEC_WARPS * 32, smem, s>>>(a);
    else spec_kernel<G, SPEC_PLAYOUT><<<grid, SPEC_WARPS * 32, smem, s>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

#endif  // GZ_SPEC_CUH
