// gz_propnet.cu — batched, bit-sliced propnet state machine for sm_100a.
//
// One warp owns a tile of 32 game states. The value of every live gate for those 32 games is one
// 32-bit word in the warp's value array in shared memory (bit g = game g). Gate programs are
// levelised 32-gate chunks (csrc/gz_devformat.h, devprog.py): lane i evaluates gate i of the chunk,
// __syncwarp() separates levels. Packed per-game rows (states, legal masks) are moved between HBM
// and shared memory with coalesced 128-bit accesses and converted to/from the bit-sliced form with
// 32x32 shuffle transposes.
//
// Replaces (batched): GGPLib::StateMachineInterface::updateBases / getLegalState / nextState /
// isTerminal / getGoalValue as called from src/cpp/puct/node.cpp:153-221,
// src/cpp/puct/evaluator.cpp:216-239 and src/cpp/pyobjects/gdltransformer_impl.cpp:50-124 of the
// reference. Results are bit-exact with oracle/propnet_ref.c (tests/test_gpu_propnet.py).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/gz_b200.h"
#include "../../include/gz_propnet_format.h"
#include "gz_devformat.h"
#include "gz_common.cuh"
#include "gz_spec.cuh"

// ------------------------------------------------------------------------------------------------
struct PnParams {
    const uint32_t* prog[3];  // A (next-state cone), LT (legal+terminal), GL (goals)
    int chunks[3];
    int words[3];
    int staged[3];
    const uint16_t* next_lit;   // [nbw*32]   padded with FALSE
    const uint16_t* legal_lit;  // [lw_used*32] padded with FALSE, role words concatenated
    const uint16_t* goal_lit;
    const uint8_t* goal_val;
    const uint32_t* init_words;  // [state_words]
    int goal_off[GZ_PN_MAX_ROLES + 1];
    int role_sizes[GZ_PN_MAX_ROLES];
    int legal_word_off[GZ_PN_MAX_ROLES + 1];
    int input_slot0[GZ_PN_MAX_ROLES];
    int n_bases, n_roles, n_inputs, state_words, legal_words, val_words, terminal_lit;
    int nbw;           // words holding base bits = ceil(n_bases/32)
    int lw_used;       // legal words actually used
    int row_stride;    // odd stride of the per-warp staging rows
    int warp_words;    // shared-memory words per warp
    int staged_words;  // shared-memory words holding staged programs
    // wide format (games with more than 32 K live slots: chess_15d, hexLG13): 32-bit literals, value
    // array in a per-CTA global scratch region (pn_wide_kernel)
    int wide;
    const uint32_t* next_lit32;
    const uint32_t* legal_lit32;
    const uint32_t* goal_lit32;
};

enum { MODE_EVAL = 0, MODE_TRANSITION = 1, MODE_PLAYOUT = 2 };
// internal goals mode: the goal cone is never evaluated (goals = 255); the PUCT engine evaluates it off
// the critical path, in a compacted pass over the terminal states only (gz_sm_eval_devcount)
#define GZ_GOALS_NEVER_INTERNAL 2

// lane i holds row i of a 32x32 bit matrix; returns column i (bit j of result = bit i of row j).
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const uint32_t m = (s == 16) ? 0x0000FFFFu : (s == 8) ? 0x00FF00FFu : (s == 4) ? 0x0F0F0F0Fu
                         : (s == 2) ? 0x33333333u : 0x55555555u;
        uint32_t y = __shfl_xor_sync(0xffffffffu, x, s);
        x = (lane & s) ? (((y >> s) & m) | (x & ~m)) : ((x & m) | ((y << s) & ~m));
    }
    return x;
}

__device__ __forceinline__ uint32_t lit_word(const uint32_t* val, uint32_t lit) {
    return val[lit >> 1] ^ (0u - (lit & 1u));
}

// Barrier of the WPT warps that share one tile (WPT == 1: just this warp).
template <int WPT>
__device__ __forceinline__ void tile_sync() {
    if (WPT == 1) __syncwarp(); else __syncthreads();
}

// Chunk offsets and level boundaries of the three programs (built at load time, build_group_tables): whoever
// evaluates only SOME chunks of a level (the warps of a shared tile, the group kernel) indexes them directly
// instead of walking the stream.
struct GroupTables {
    const uint32_t* desc[3];     // per chunk: word offset into the program | literal rows << 28
    const int* level_first[3];   // per level: index of its first chunk; [n_levels] = n_chunks
    int n_levels[3];
};
#define GZ_GROUP_TABLES_DEFINED 1

// ---- bulk staging (TMA, non-tensor form): one thread issues cp.async.bulk global -> shared for a whole gate
// program and everybody waits on an mbarrier transaction count (SASS: UBLKCP + SYNCS), instead of 256 threads
// copying it word by word through registers
__device__ __forceinline__ void bulk_bar_init(uint64_t* bar) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_bar_expect(uint64_t* bar, uint32_t bytes) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gmem_src), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void bulk_bar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    }
}

// Evaluates one chunk stream on the tile's value array. With WPT > 1 the CTA's WPT warps share the
// tile: warp `wt` evaluates chunks wt, wt+WPT, ... and the CTA meets at every level boundary.
template <int WPT>
__device__ __forceinline__ void run_program(const uint32_t* __restrict__ prog, int n_chunks,
                                            uint32_t* __restrict__ val, int lane, int wt) {
    const uint32_t* p = prog;
    for (int c = 0; c < n_chunks; ++c) {
        const uint32_t hdr = p[0];
        const int nl2 = (int)((hdr & 0xffu) >> 1);
        if (WPT == 1 || (c % WPT) == wt) {
            const uint32_t ow = p[1 + (lane >> 1)];
            const uint32_t out = (lane & 1) ? (ow >> 16) : (ow & 0xffffu);
            uint32_t acc = 0xffffffffu;
            const uint32_t* lp = p + 17 + lane;
#pragma unroll 2
            for (int k = 0; k < nl2; ++k) {
                const uint32_t pr = lp[k * 32];
                acc &= lit_word(val, pr & 0xffffu) & lit_word(val, pr >> 16);
            }
            val[out] = acc;
        }
        p += 17 + 32 * nl2;
        if (hdr & 0x100u) tile_sync<WPT>();
    }
}

// The same evaluation for a tile shared by WPT warps, driven by the chunk / level tables: warp wt takes chunks
// wt, wt+WPT, .. of every level and nobody walks the headers of chunks it does not evaluate (in the PUCT step,
// where a tile has a CTA to itself, that dependent walk was a third of the kernel's stall samples).
template <int WPT>
__device__ __forceinline__ void run_program_levels(const uint32_t* __restrict__ prog, const uint32_t* __restrict__ desc,
                                                   const int* __restrict__ level_first, int n_levels,
                                                   uint32_t* __restrict__ val, int lane, int wt) {
    // level bounds and this warp's first chunk descriptor are fetched ONE LEVEL AHEAD: on the critical path they
    // were a dependent global load per level right after the barrier (47 % of the kernel's stall samples)
    int c_lo = __ldg(level_first);
    int c_hi = n_levels > 0 ? __ldg(level_first + 1) : c_lo;
    uint32_t d_first = (c_lo + wt < c_hi) ? __ldg(desc + c_lo + wt) : 0u;
    for (int l = 0; l < n_levels; ++l) {
        const int c_hi2 = (l + 2 <= n_levels) ? __ldg(level_first + l + 2) : c_hi;
        // (the next level's first descriptor needs c_hi, known since the previous iteration)
        const uint32_t d_next = (l + 1 < n_levels && c_hi + wt < c_hi2) ? __ldg(desc + c_hi + wt) : 0u;
        for (int c = c_lo + wt; c < c_hi; c += WPT) {
            const uint32_t d = (c == c_lo + wt) ? d_first : __ldg(desc + c);
            const uint32_t* p = prog + (d & 0x0fffffffu);
            const int nl2 = (int)(d >> 28) >> 1;
            const uint32_t ow = p[1 + (lane >> 1)];
            const uint32_t out = (lane & 1) ? (ow >> 16) : (ow & 0xffffu);
            uint32_t acc = 0xffffffffu;
            const uint32_t* lp = p + 17 + lane;
#pragma unroll 2
            for (int k = 0; k < nl2; ++k) {
                const uint32_t pr = lp[k * 32];
                acc &= lit_word(val, pr & 0xffffu) & lit_word(val, pr >> 16);
            }
            val[out] = acc;
        }
        c_lo = c_hi;
        c_hi = c_hi2;
        d_first = d_next;
        tile_sync<WPT>();
    }
}

// Coalesced copy of `rows` packed rows (row_words words each, multiple of 4) HBM -> staging rows.
__device__ __forceinline__ void load_rows(const uint32_t* __restrict__ src, int row_words, int rows,
                                          uint32_t* stage, int rs, int lane) {
    if (lane >= rows)
        for (int c = 0; c < row_words; ++c) stage[lane * rs + c] = 0;
    const int qpr = row_words >> 2;
    const int nq = rows * qpr;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    for (int q = lane; q < nq; q += 32) {
        const uint4 v = s4[q];
        const int r = q / qpr;
        const int c = (q - r * qpr) << 2;
        uint32_t* d = stage + r * rs + c;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncwarp();
}

__device__ __forceinline__ void store_rows(uint32_t* __restrict__ dst, int row_words, int rows,
                                           const uint32_t* stage, int rs, int lane) {
    __syncwarp();
    const int qpr = row_words >> 2;
    const int nq = rows * qpr;
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (int q = lane; q < nq; q += 32) {
        const int r = q / qpr;
        const int c = (q - r * qpr) << 2;
        const uint32_t* s = stage + r * rs + c;
        d4[q] = make_uint4(s[0], s[1], s[2], s[3]);
    }
    __syncwarp();
}

template <int MODE, int WPT>
__global__ void __launch_bounds__(1024, 1)
pn_kernel(const __grid_constant__ PnParams P, const __grid_constant__ GroupTables GT, int64_t n, int64_t slot0,
          const uint32_t* states_in, const int16_t* __restrict__ moves_in,
          uint32_t* states_out, uint32_t* legal,
          uint8_t* terminal, uint8_t* __restrict__ goals, int16_t* __restrict__ moves_out,
          unsigned long long* __restrict__ transitions, uint64_t seed, uint64_t step, int goals_mode,
          const int* __restrict__ n_dev) {
    extern __shared__ __align__(16) uint32_t smem[];
    if (n_dev) n = (int64_t)*n_dev;      // batch size decided on the device (compacted sub-batches)

    // ---- stage gate programs in shared memory (those the host decided fit): one bulk copy each, in flight while
    // the first tile's rows are loaded and transposed; everybody waits on the barrier before the first program runs
    __shared__ __align__(8) uint64_t stage_bar;
    const uint32_t* prog[3];
    bool staging = false;
    {
        uint32_t* sp = smem;
        uint32_t bytes[3];
        uint32_t total = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            bytes[k] = 0;
            if (P.staged[k]) {
                bytes[k] = (uint32_t)((P.words[k] + 3) & ~3) * 4u;    // (the device copy is padded to 16 bytes)
                total += bytes[k];
                prog[k] = sp;
                sp += (P.words[k] + 3) & ~3;
            } else {
                prog[k] = P.prog[k];
            }
        }
        staging = total != 0 && WPT > 1;
        if (WPT == 1 && total != 0) {
            // a persistent CTA of many warps with a tile each: plain cooperative copy (measured: the bulk-copy path
            // costs this configuration 5 % — reversi 573 -> 534 M transitions/s — for no gain, staging happens once)
            uint32_t* dst = smem;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (P.staged[k]) {
                    for (int i = threadIdx.x; i < P.words[k]; i += blockDim.x) dst[i] = P.prog[k][i];
                    dst += (P.words[k] + 3) & ~3;
                }
            }
            __syncthreads();
        }
        if (staging) {
            if (threadIdx.x == 0) bulk_bar_init(&stage_bar);
            __syncthreads();
            if (threadIdx.x == 0) {
                bulk_bar_expect(&stage_bar, total);
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (bytes[k]) bulk_copy_g2s(const_cast<uint32_t*>(prog[k]), P.prog[k], bytes[k], &stage_bar);
            }
        }
    }
    if (WPT > 1) {
        // the chunk / level tables are read through the read-only path: touch them once so that the per-level
        // look-ups hit in L1
        for (int k = 0; k < 3; ++k) {
            for (int i = threadIdx.x * 32; i < P.chunks[k]; i += blockDim.x * 32)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(GT.desc[k] + i));
            for (int i = threadIdx.x * 32; i <= GT.n_levels[k]; i += blockDim.x * 32)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(GT.level_first[k] + i));
        }
    }

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // WPT == 1: every warp owns its own tile (and value array). WPT > 1: the CTA is WPT warps sharing
    // ONE tile; warp 0 does the row I/O and transposes, all warps evaluate the gate programs.
    const int warps = WPT == 1 ? (int)(blockDim.x >> 5) : 1;
    const int wt = WPT == 1 ? 0 : warp;
    const bool io = (WPT == 1) || (warp == 0);
    uint32_t* val = smem + P.staged_words + (WPT == 1 ? warp : 0) * P.warp_words;
    uint32_t* stage = val + P.val_words;
    uint32_t* slices = stage + 32 * P.row_stride;
    const int rs = P.row_stride;
    const int R = P.n_roles;
    if (lane == 0 && io) val[0] = 0xffffffffu;

    const int64_t n_tiles = (n + 31) >> 5;
    for (int64_t tile = (int64_t)blockIdx.x * warps + (WPT == 1 ? warp : 0); tile < n_tiles;
         tile += (int64_t)gridDim.x * warps) {
        const int64_t g0 = tile << 5;
        const int64_t game = g0 + lane;
        const int rows = (int)((n - g0) < 32 ? (n - g0) : 32);
        const bool valid = lane < rows;
        const uint32_t valid_mask = rows == 32 ? 0xffffffffu : ((1u << rows) - 1u);
        uint32_t reset_mask = 0, pass_mask = 0;
        int mv[GZ_PN_MAX_ROLES];

        if (io) {
        if (MODE == MODE_PLAYOUT) {
            // ---- pick one uniformly random legal action per role from the stored legal rows
            const bool term = valid ? (terminal[game] != 0) : false;
            reset_mask = __ballot_sync(0xffffffffu, term);
            load_rows(legal + g0 * P.legal_words, P.legal_words, rows, stage, rs, lane);
#pragma unroll
            for (int r = 0; r < GZ_PN_MAX_ROLES; ++r) {
                mv[r] = -1;
                if (r < R && valid && !term) {
                    const uint32_t* w = stage + lane * rs + P.legal_word_off[r];
                    const int nw = P.legal_word_off[r + 1] - P.legal_word_off[r];
                    uint32_t cnt = 0;
                    for (int j = 0; j < nw; ++j) cnt += __popc(w[j]);
                    uint32_t pick = gz_pick(gz_rand32(seed, (uint64_t)(slot0 + game), step, (uint32_t)r), cnt);
                    int idx = 0;
                    for (int j = 0; j < nw; ++j) {
                        uint32_t x = w[j];
                        const uint32_t c = __popc(x);
                        if (pick < c) {
                            for (uint32_t i = 0; i < pick; ++i) x &= x - 1;
                            idx = j * 32 + (__ffs(x) - 1);
                            break;
                        }
                        pick -= c;
                    }
                    mv[r] = idx;
                }
                if (r < R && valid && moves_out) moves_out[game * R + r] = (int16_t)mv[r];
            }
            __syncwarp();
        } else if (MODE == MODE_TRANSITION) {
#pragma unroll
            for (int r = 0; r < GZ_PN_MAX_ROLES; ++r)
                mv[r] = (r < R && valid) ? (int)moves_in[game * R + r] : -1;
            pass_mask = __ballot_sync(0xffffffffu, valid && mv[0] < 0);
        }

        // ---- load packed states, slice them into the base slots
        load_rows(states_in + g0 * P.state_words, P.state_words, rows, stage, rs, lane);
        for (int j = 0; j < P.nbw; ++j) {
            const uint32_t t = warp_transpose32(stage[lane * rs + j], lane);
            const int b = j * 32 + lane;
            if (b < P.n_bases) val[1 + b] = t;
        }

        if (MODE != MODE_EVAL) {
            // ---- one-hot the chosen actions into the input slots
            for (int i = lane; i < P.n_inputs; i += 32) val[1 + P.n_bases + i] = 0;
            __syncwarp();
#pragma unroll
            for (int r = 0; r < GZ_PN_MAX_ROLES; ++r)
                if (r < R && mv[r] >= 0) atomicOr(&val[P.input_slot0[r] + mv[r]], 1u << lane);
        }
        }   // io
        tile_sync<WPT>();

        if (staging) { bulk_bar_wait(&stage_bar, 0); staging = false; }    // (first tile only: phase 0 completes once)
        if (MODE != MODE_EVAL) {
            // ---- next-state cone
            if (WPT > 1) run_program_levels<WPT>(prog[0], GT.desc[0], GT.level_first[0], GT.n_levels[0], val, lane, wt);
            else run_program<WPT>(prog[0], P.chunks[0], val, lane, wt);
            if (io) {
            for (int j = 0; j < P.nbw; ++j) {
                const int b = j * 32 + lane;
                uint32_t t = lit_word(val, P.next_lit[b]);
                if (MODE == MODE_PLAYOUT) {
                    const uint32_t ib = (P.init_words[j] >> lane) & 1u;
                    t = (t & ~reset_mask) | (ib ? reset_mask : 0u);
                } else {
                    // games whose first move index is negative keep their state (identity): used by
                    // the PUCT pipeline to evaluate a root without moving
                    if (b < P.n_bases) t = (t & ~pass_mask) | (val[1 + b] & pass_mask);
                }
                t &= valid_mask;
                slices[b] = t;
                stage[lane * rs + j] = warp_transpose32(t, lane);
            }
            for (int j = P.nbw; j < P.state_words; ++j) stage[lane * rs + j] = 0;
            store_rows(states_out + g0 * P.state_words, P.state_words, rows, stage, rs, lane);
            for (int j = 0; j < P.nbw; ++j) {
                const int b = j * 32 + lane;
                if (b < P.n_bases) val[1 + b] = slices[b];
            }
            }   // io
        }
        tile_sync<WPT>();

        // ---- views of the (new) state: legal masks, terminal, goals
        if (WPT > 1) run_program_levels<WPT>(prog[1], GT.desc[1], GT.level_first[1], GT.n_levels[1], val, lane, wt);
        else run_program<WPT>(prog[1], P.chunks[1], val, lane, wt);
        const uint32_t tw = lit_word(val, (uint32_t)P.terminal_lit) & valid_mask;
        const bool need_goals = (goals_mode == GZ_GOALS_ALWAYS) || (goals_mode == GZ_GOALS_TERMINAL_ONLY && tw != 0u);
        if (need_goals) {
            if (WPT > 1) run_program_levels<WPT>(prog[2], GT.desc[2], GT.level_first[2], GT.n_levels[2], val, lane, wt);
            else run_program<WPT>(prog[2], P.chunks[2], val, lane, wt);
        }

        if (io) {
        for (int j = 0; j < P.lw_used; ++j) {
            const uint32_t t = lit_word(val, P.legal_lit[j * 32 + lane]) & valid_mask;
            stage[lane * rs + j] = warp_transpose32(t, lane);
        }
        for (int j = P.lw_used; j < P.legal_words; ++j) stage[lane * rs + j] = 0;
        store_rows(legal + g0 * P.legal_words, P.legal_words, rows, stage, rs, lane);

        const bool is_term = (tw >> lane) & 1u;
        if (valid) terminal[game] = is_term ? 1 : 0;
        for (int r = 0; r < R; ++r) {
            uint32_t gv = 255u;
            if (need_goals) {
                for (int k = P.goal_off[r]; k < P.goal_off[r + 1]; ++k) {
                    const uint32_t w = lit_word(val, P.goal_lit[k]);
                    if ((w >> lane) & 1u) gv = P.goal_val[k];
                }
            }
            if (goals_mode == GZ_GOALS_TERMINAL_ONLY && !is_term) gv = 255u;
            if (valid) goals[game * R + r] = (uint8_t)gv;
        }
        if (MODE == MODE_PLAYOUT && transitions && lane == 0) {
            const int cnt = __popc(valid_mask & ~reset_mask);
            if (cnt) atomicAdd(transitions, (unsigned long long)cnt);
        }
        }   // io
        tile_sync<WPT>();
    }
    if (staging) bulk_bar_wait(&stage_bar, 0);     // a warp without a tile: the copy into this CTA's memory must land first
}

// ------------------------------------------------------------------------------------------------
// Wide games. The value array of a tile (55 K words for chess_15d, 412 K for hexLG13) does not fit in
// shared memory, so it lives in a per-CTA region of a global scratch buffer that stays L2-resident;
// one CTA of WIDE_WARPS warps owns one tile, warp w evaluates chunks w, w+WIDE_WARPS, ... of a level
// and the CTA meets at level boundaries. Same bit-sliced evaluation and the same results as pn_kernel.
#define WIDE_WARPS 16

__device__ __forceinline__ uint32_t wlit(const uint32_t* val, uint32_t lit) {
    return __ldcg(val + (lit >> 1)) ^ (0u - (lit & 1u));
}

__device__ __forceinline__ void run_program_wide(const uint32_t* __restrict__ prog, int n_chunks,
                                                 uint32_t* val, int lane, int warp) {
    const uint32_t* p = prog;
    for (int c = 0; c < n_chunks; ++c) {
        const uint32_t hdr = __ldg(p);
        const int nl = (int)(hdr & 0xffu);
        if ((c % WIDE_WARPS) == warp) {
            const uint32_t out = __ldg(p + 1 + lane);
            uint32_t acc = 0xffffffffu;
            const uint32_t* lp = p + 33 + lane;
#pragma unroll 4
            for (int k = 0; k < nl; ++k) acc &= wlit(val, __ldg(lp + k * 32));
            __stcg(val + out, acc);
        }
        p += 33 + 32 * nl;
        if (hdr & 0x100u) __syncthreads();
    }
}

template <int MODE>
__global__ void __launch_bounds__(WIDE_WARPS * 32)
pn_wide_kernel(const __grid_constant__ PnParams P, int64_t n, int64_t slot0,
               const uint32_t* states_in, const int16_t* __restrict__ moves_in,
               uint32_t* states_out, uint32_t* legal, uint8_t* terminal, uint8_t* __restrict__ goals,
               int16_t* __restrict__ moves_out, unsigned long long* __restrict__ transitions,
               uint64_t seed, uint64_t step, int goals_mode, uint32_t* scratch, int64_t scratch_stride,
               const int* __restrict__ n_dev) {
    __shared__ int mv_sm[32][GZ_PN_MAX_ROLES];
    if (n_dev) n = (int64_t)*n_dev;
    __shared__ uint32_t masks_sm[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* val = scratch + (int64_t)blockIdx.x * scratch_stride;
    uint32_t* slices = val + P.val_words;          // [nbw * 32] next-state slices of the tile
    const int R = P.n_roles;
    const int SW = P.state_words, LW = P.legal_words;
    if (threadIdx.x == 0) __stcg(val, 0xffffffffu);
    const int64_t n_tiles = (n + 31) >> 5;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t g0 = tile << 5;
        const int64_t game = g0 + lane;
        const int rows = (int)((n - g0) < 32 ? (n - g0) : 32);
        const bool valid = lane < rows;
        const uint32_t valid_mask = rows == 32 ? 0xffffffffu : ((1u << rows) - 1u);

        // ---- moves (warp 0): chosen from the stored legal rows (PLAYOUT) or given (TRANSITION)
        if (warp == 0) {
            uint32_t reset_mask = 0, pass_mask = 0;
            int mv[GZ_PN_MAX_ROLES];
            if (MODE == MODE_PLAYOUT) {
                const bool term = valid ? (terminal[game] != 0) : false;
                reset_mask = __ballot_sync(0xffffffffu, term);
                for (int r = 0; r < R; ++r) {
                    mv[r] = -1;
                    if (valid && !term) {
                        const uint32_t* w = legal + game * LW + P.legal_word_off[r];
                        const int nw = P.legal_word_off[r + 1] - P.legal_word_off[r];
                        uint32_t cnt = 0;
                        for (int j = 0; j < nw; ++j) cnt += __popc(w[j]);
                        uint32_t pick = gz_pick(gz_rand32(seed, (uint64_t)(slot0 + game), step, (uint32_t)r), cnt);
                        int idx = 0;
                        for (int j = 0; j < nw; ++j) {
                            uint32_t x = w[j];
                            const uint32_t c = __popc(x);
                            if (pick < c) {
                                for (uint32_t i = 0; i < pick; ++i) x &= x - 1;
                                idx = j * 32 + (__ffs(x) - 1);
                                break;
                            }
                            pick -= c;
                        }
                        mv[r] = idx;
                    }
                    if (valid && moves_out) moves_out[game * R + r] = (int16_t)mv[r];
                }
            } else if (MODE == MODE_TRANSITION) {
                for (int r = 0; r < R; ++r) mv[r] = valid ? (int)moves_in[game * R + r] : -1;
                pass_mask = __ballot_sync(0xffffffffu, valid && mv[0] < 0);
            }
            if (MODE != MODE_EVAL)
                for (int r = 0; r < R; ++r) mv_sm[lane][r] = mv[r];
            if (lane == 0) { masks_sm[0] = reset_mask; masks_sm[1] = pass_mask; }
        }
        // ---- packed states -> base slots (every warp takes state words w, w+WIDE_WARPS, ...)
        for (int j = warp; j < P.nbw; j += WIDE_WARPS) {
            const uint32_t x = valid ? states_in[game * SW + j] : 0u;
            const uint32_t t = warp_transpose32(x, lane);
            const int b = j * 32 + lane;
            if (b < P.n_bases) __stcg(val + 1 + b, t);
        }
        if (MODE != MODE_EVAL)
            for (int i = threadIdx.x; i < P.n_inputs; i += blockDim.x) __stcg(val + 1 + P.n_bases + i, 0u);
        __syncthreads();
        const uint32_t reset_mask = masks_sm[0], pass_mask = masks_sm[1];

        if (MODE != MODE_EVAL) {
            if (warp == 0)
                for (int r = 0; r < R; ++r) {
                    const int m = mv_sm[lane][r];
                    if (m >= 0) atomicOr(val + P.input_slot0[r] + m, 1u << lane);
                }
            __syncthreads();
            // ---- next-state cone
            run_program_wide(P.prog[0], P.chunks[0], val, lane, warp);
            __syncthreads();
            for (int j = warp; j < P.nbw; j += WIDE_WARPS) {
                const int b = j * 32 + lane;
                uint32_t t = wlit(val, __ldg(P.next_lit32 + b));
                if (MODE == MODE_PLAYOUT) {
                    const uint32_t ib = (P.init_words[j] >> lane) & 1u;
                    t = (t & ~reset_mask) | (ib ? reset_mask : 0u);
                } else if (b < P.n_bases) {
                    t = (t & ~pass_mask) | (__ldcg(val + 1 + b) & pass_mask);
                }
                t &= valid_mask;
                slices[b] = t;
                const uint32_t x = warp_transpose32(t, lane);
                if (valid) states_out[game * SW + j] = x;
            }
            for (int j = P.nbw + warp; j < SW; j += WIDE_WARPS)
                if (valid) states_out[game * SW + j] = 0u;
            __syncthreads();
            for (int b = threadIdx.x; b < P.n_bases; b += blockDim.x) __stcg(val + 1 + b, slices[b]);
            __syncthreads();
        }

        // ---- views of the (new) state
        run_program_wide(P.prog[1], P.chunks[1], val, lane, warp);
        __syncthreads();
        const uint32_t tw = wlit(val, (uint32_t)P.terminal_lit) & valid_mask;
        const bool need_goals = (goals_mode == GZ_GOALS_ALWAYS) || (goals_mode == GZ_GOALS_TERMINAL_ONLY && tw != 0u);
        if (need_goals) { run_program_wide(P.prog[2], P.chunks[2], val, lane, warp); __syncthreads(); }

        for (int j = warp; j < LW; j += WIDE_WARPS) {
            uint32_t x = 0u;
            if (j < P.lw_used) {
                const uint32_t t = wlit(val, __ldg(P.legal_lit32 + j * 32 + lane)) & valid_mask;
                x = warp_transpose32(t, lane);
            }
            if (valid) legal[game * LW + j] = x;
        }
        if (warp == 0) {
            const bool is_term = (tw >> lane) & 1u;
            if (valid) terminal[game] = is_term ? 1 : 0;
            for (int r = 0; r < R; ++r) {
                uint32_t gv = 255u;
                if (need_goals) {
                    for (int k = P.goal_off[r]; k < P.goal_off[r + 1]; ++k) {
                        const uint32_t w = wlit(val, __ldg(P.goal_lit32 + k));
                        if ((w >> lane) & 1u) gv = P.goal_val[k];
                    }
                }
                if (goals_mode == GZ_GOALS_TERMINAL_ONLY && !is_term) gv = 255u;
                if (valid) goals[game * R + r] = (uint8_t)gv;
            }
            if (MODE == MODE_PLAYOUT && transitions && lane == 0) {
                const int cnt = __popc(valid_mask & ~reset_mask);
                if (cnt) atomicAdd(transitions, (unsigned long long)cnt);
            }
        }
        __syncthreads();
    }
}

#include "gz_propnet_group.cuh"

// ------------------------------------------------------------------------------------------------
struct gz_propnet {
    int device = 0;
    PnParams P{};
    // group kernel (gz_propnet_group.cuh): chunk / level tables, cluster size, scratch per launching stream
    GroupTables GT{};
    void* d_group_tables = nullptr;
    int group_cl = 1;                    // CTAs per cluster: 1 (narrow levels) or 8 (wide levels)
    bool group_cl16_ok = false;          // clusters of 16 CTAs can be launched (non-portable size, opt-in per kernel)
    int group_clusters = 0;              // clusters resident at once
    int64_t group_stride = 0;            // scratch words per cluster
    int64_t group_min_batch = GROUP_GAMES;
    std::vector<std::pair<cudaStream_t, uint32_t*>> group_scratch;
    void* d_arena = nullptr;
    int warps_per_cta = 0;
    int grid_max = 0;
    size_t smem_bytes = 0;
    int n_sms = 0;
    // scratch for the host-buffer API
    void* d_scratch = nullptr;
    size_t scratch_bytes = 0;
    cudaStream_t host_stream = nullptr;
    // wide games: one value-array scratch per stream that launches (engines on different streams may
    // share one gz_propnet)
    // game-specialised thread-sliced kernels (gz_spec.cuh), used for batches >= spec_min_batch
    const GzSpecEntry* spec = nullptr;
    int64_t spec_min_batch = (int64_t)1 << 18;
    std::vector<std::pair<cudaStream_t, uint32_t*>> wide_scratch;
    int wide_grid = 0;
    int64_t wide_stride = 0;
    // deferred goal pass (deferred_goals): compacted copies of the terminal states of a batch, per launching stream
    struct GoalScratch { cudaStream_t st; int64_t cap; void* mem; };
    std::vector<GoalScratch> goal_scratch;
    int64_t deferred_goals_min_batch = 4096;
};

extern "C" int gz_version(void) { return 100; }

// registry of specialised games (filled by static initialisers of csrc/gz_spec_<game>.cu)
static GzSpecEntry* g_spec_head = nullptr;
void gz_spec_register(GzSpecEntry* e) { e->next = g_spec_head; g_spec_head = e; }
const GzSpecEntry* gz_spec_find(uint64_t hash) {
    for (const GzSpecEntry* e = g_spec_head; e; e = e->next) if (e->hash == hash) return e;
    return nullptr;
}
static uint64_t fnv1a64(const void* data, size_t bytes) {
    const uint8_t* p = (const uint8_t*)data;
    uint64_t h = 0xcbf29ce484222325ULL;
    for (size_t i = 0; i < bytes; ++i) { h ^= p[i]; h *= 0x100000001b3ULL; }
    return h;
}

static int configure_launch(gz_propnet* pn, const int32_t* h) {
    PnParams& P = pn->P;
    cudaDeviceProp prop;
    GZ_CUDA(cudaGetDeviceProperties(&prop, pn->device));
    pn->n_sms = prop.multiProcessorCount;
    if (P.wide) {
        // value array + next-state slices per CTA in global scratch (allocated per launching stream)
        pn->wide_grid = pn->n_sms * 2;
        pn->wide_stride = (((int64_t)P.val_words + (int64_t)P.nbw * 32) + 31) & ~(int64_t)31;
        pn->warps_per_cta = WIDE_WARPS;
        pn->grid_max = pn->wide_grid;
        pn->smem_bytes = 0;
        P.row_stride = P.warp_words = P.staged_words = 0;
        for (int k = 0; k < 3; ++k) P.staged[k] = 0;
        return GZ_OK;
    }
    const size_t max_smem = prop.sharedMemPerBlockOptin;  // 227 KB on sm_100
    const int rs = std::max(P.state_words, P.legal_words) | 1;
    P.row_stride = rs;
    P.warp_words = ((P.val_words + 32 * rs + P.nbw * 32) + 3) & ~3;
    const size_t warp_bytes = (size_t)P.warp_words * 4;
    // stage programs (A, LT, GL in that priority) while at least MIN_WARPS warps still fit
    int MIN_WARPS = 8;
    { const char* mw = getenv("GZ_PN_MIN_WARPS"); if (mw && atoi(mw) > 0) MIN_WARPS = atoi(mw); }   // measurement switch
    size_t staged = 0;
    for (int k = 0; k < 3; ++k) {
        const size_t w = (size_t)((P.words[k] + 3) & ~3) * 4;
        if (P.words[k] > 0 && staged + w + MIN_WARPS * warp_bytes <= max_smem) {
            P.staged[k] = 1;
            staged += w;
        } else {
            P.staged[k] = 0;
        }
    }
    P.staged_words = (int)(staged / 4);
    int warps = (int)((max_smem - staged) / warp_bytes);
    if (warps < 1) {
        gz_set_error("propnet value array (%zu bytes per warp) does not fit in shared memory", warp_bytes);
        return GZ_ERR_LIMIT;
    }
    warps = std::min(warps, 32);
    pn->warps_per_cta = warps;
    pn->smem_bytes = staged + (size_t)warps * warp_bytes;
    pn->grid_max = pn->n_sms;  // one persistent CTA per SM
#define GZ_SET_SMEM(M, W) GZ_CUDA(cudaFuncSetAttribute(pn_kernel<M, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pn->smem_bytes))
    GZ_SET_SMEM(MODE_EVAL, 1); GZ_SET_SMEM(MODE_TRANSITION, 1); GZ_SET_SMEM(MODE_PLAYOUT, 1);
    GZ_SET_SMEM(MODE_EVAL, 4); GZ_SET_SMEM(MODE_TRANSITION, 4); GZ_SET_SMEM(MODE_PLAYOUT, 4);
    GZ_SET_SMEM(MODE_EVAL, 8); GZ_SET_SMEM(MODE_TRANSITION, 8); GZ_SET_SMEM(MODE_PLAYOUT, 8);
#undef GZ_SET_SMEM
    (void)h;
    return GZ_OK;
}

// Chunk offsets and level boundaries of the three programs, for the group kernel (nobody walks the stream there).
static int build_group_tables(gz_propnet* pn, const int32_t* h) {
    const PnParams& P = pn->P;
    const int offs[3] = {D_OFF_PROG_A, D_OFF_PROG_LT, D_OFF_PROG_GL};
    std::vector<uint32_t> desc[3];
    std::vector<int> lf[3];
    int max_level_chunks = 0;
    for (int k = 0; k < 3; ++k) {
        const uint32_t* w = (const uint32_t*)(h + h[offs[k]]);
        size_t p = 0;
        lf[k].push_back(0);
        for (int c = 0; c < P.chunks[k]; ++c) {
            const uint32_t hdr = w[p];
            const int nl = P.wide ? (int)(hdr & 0xffu) : (int)(hdr & 0xffu);          // literal rows (narrow: even)
            if (p >= (1u << 28) || nl > 8) { gz_set_error("program outside the group kernel's limits (%d literal rows)", nl); return GZ_ERR_LIMIT; }
            desc[k].push_back((uint32_t)p | ((uint32_t)nl << 28));
            p += P.wide ? (33 + 32 * (size_t)nl) : (17 + 32 * (size_t)(nl >> 1));
            if (hdr & 0x100u) {
                max_level_chunks = std::max(max_level_chunks, c + 1 - lf[k].back());
                lf[k].push_back(c + 1);
            }
        }
        if (lf[k].back() != P.chunks[k]) lf[k].push_back(P.chunks[k]);            // (a stream always ends on a level)
        if (desc[k].empty()) desc[k].push_back(0);
    }
    auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
    size_t pos = 0, o_desc[3], o_lf[3];
    for (int k = 0; k < 3; ++k) { o_desc[k] = pos; pos += al(desc[k].size() * 4); o_lf[k] = pos; pos += al(lf[k].size() * 4); }
    std::vector<uint8_t> buf(pos, 0);
    for (int k = 0; k < 3; ++k) {
        memcpy(buf.data() + o_desc[k], desc[k].data(), desc[k].size() * 4);
        memcpy(buf.data() + o_lf[k], lf[k].data(), lf[k].size() * 4);
    }
    GZ_CUDA(cudaMalloc(&pn->d_group_tables, pos));
    GZ_CUDA(cudaMemcpy(pn->d_group_tables, buf.data(), pos, cudaMemcpyHostToDevice));
    for (int k = 0; k < 3; ++k) {
        pn->GT.desc[k] = (const uint32_t*)((const uint8_t*)pn->d_group_tables + o_desc[k]);
        pn->GT.level_first[k] = (const int*)((const uint8_t*)pn->d_group_tables + o_lf[k]);
        pn->GT.n_levels[k] = (int)lf[k].size() - 1;
    }
    // levels wider than a CTA's 16 warps can use: spread them over a cluster of 8 CTAs
    pn->group_cl = max_level_chunks > 4 * GROUP_WARPS ? 8 : 1;
    // Measured (profiles/r02_propnet_games.md): the group kernel pays two L2 round trips per level, the per-tile
    // kernels a shared-memory access. Programs with few, wide levels (chess_15d: 34 levels of up to 300 chunks,
    // hexLG13: 19 of up to 1400) run 2.5-3x faster in groups; deep narrow ones (reversi: 175 levels of 4 chunks)
    // stay with the per-tile kernels unless asked (gz_propnet_set_group_min_batch) — and so does every narrow-format
    // game, whatever its widest level: reversi's goal cone has one level of more than 64 chunks, yet its group
    // kernel is 15x slower than pn_kernel (r2i_pn_reversi_*: 10 M vs 158 M transitions/s; amazons 24 M vs 44 M,
    // connect6 92 M vs 685 M), and inside the PUCT step it cost 1.23 ms against 0.31 ms.
    pn->group_min_batch = (P.wide && pn->group_cl == 8) ? GROUP_GAMES : -1;
    if (P.wide && pn->group_cl == 8) {
        const char* off = getenv("GZ_GROUP_NO_CL16");        // measurement switch
        bool ok = !(off && off[0] == '1');
        ok = ok && cudaFuncSetAttribute(pn_group_kernel<MODE_EVAL, true, 16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(pn_group_kernel<MODE_TRANSITION, true, 16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(pn_group_kernel<MODE_PLAYOUT, true, 16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
        cudaGetLastError();
        pn->group_cl16_ok = ok;
    }
    pn->group_clusters = pn->group_cl == 8 ? pn->n_sms / 8 : pn->n_sms;
    { const char* gc = getenv("GZ_GROUP_CLUSTERS"); if (gc && atoi(gc) > 0) pn->group_clusters = std::min(pn->group_clusters, atoi(gc)); }   // measurement switch
    const int64_t slices = std::max<int64_t>((int64_t)P.nbw * 32 * GROUP_TILES, (int64_t)GROUP_TILES * GZ_PN_MAX_ROLES * 32);
    pn->group_stride = (((int64_t)P.val_words * GROUP_TILES + slices + 2 * GROUP_TILES) + 31) & ~(int64_t)31;
    return GZ_OK;
}

extern "C" int gz_propnet_create(const int32_t* h, int64_t n_words, int device, gz_propnet** out) {
    if (!h || !out) { gz_set_error("null argument"); return GZ_ERR_ARG; }
    *out = nullptr;
    if (n_words < GZ_DEV_HEADER_WORDS || h[D_MAGIC] != GZ_DEV_MAGIC || h[D_VERSION] != GZ_DEV_VERSION ||
        h[D_TOTAL_WORDS] != n_words) {
        gz_set_error("not a gz device program blob (magic/version/size mismatch)");
        return GZ_ERR_FORMAT;
    }
    const int R = h[D_N_ROLES];
    if (R < 1 || R > GZ_PN_MAX_ROLES) { gz_set_error("unsupported role count %d", R); return GZ_ERR_LIMIT; }
    GZ_CUDA(cudaSetDevice(device));
    gz_propnet* pn = new gz_propnet();
    pn->device = device;
    PnParams& P = pn->P;
    P.n_bases = h[D_N_BASES];
    P.n_roles = R;
    P.n_inputs = h[D_N_INPUTS];
    P.state_words = h[D_STATE_WORDS];
    P.legal_words = h[D_LEGAL_WORDS];
    P.val_words = h[D_VAL_WORDS];
    P.terminal_lit = h[D_TERMINAL_LIT];
    P.nbw = (P.n_bases + 31) / 32;
    P.wide = h[D_WIDE] ? 1 : 0;
    const int32_t* role_sizes = h + h[D_OFF_ROLE_SIZES];
    const int32_t* lwo = h + h[D_OFF_LEGAL_WORD_OFF];
    const int32_t* goal_off = h + h[D_OFF_GOAL_OFF];
    int in_off = 1 + P.n_bases;
    for (int r = 0; r < R; ++r) {
        P.role_sizes[r] = role_sizes[r];
        P.input_slot0[r] = in_off;
        in_off += role_sizes[r];
    }
    for (int r = 0; r <= R; ++r) { P.legal_word_off[r] = lwo[r]; P.goal_off[r] = goal_off[r]; }
    P.lw_used = lwo[R];

    // ---- host-side tables: padded literal arrays
    std::vector<uint16_t> next_lit((size_t)P.nbw * 32, 1);
    for (int b = 0; b < P.n_bases; ++b) next_lit[b] = (uint16_t)h[h[D_OFF_NEXT_LIT] + b];
    std::vector<uint16_t> legal_lit((size_t)P.lw_used * 32, 1);
    {
        int src = h[D_OFF_LEGAL_LIT];
        for (int r = 0; r < R; ++r)
            for (int a = 0; a < role_sizes[r]; ++a) legal_lit[(size_t)lwo[r] * 32 + a] = (uint16_t)h[src++];
    }
    const int n_goal = goal_off[R];
    std::vector<uint16_t> goal_lit(std::max(n_goal, 1));
    std::vector<uint8_t> goal_val(std::max(n_goal, 1));
    for (int k = 0; k < n_goal; ++k) {
        goal_lit[k] = (uint16_t)h[h[D_OFF_GOAL_LIT] + k];
        goal_val[k] = (uint8_t)h[h[D_OFF_GOAL_VAL] + k];
    }
    const int offs[3] = {D_OFF_PROG_A, D_OFF_PROG_LT, D_OFF_PROG_GL};
    const int chs[3] = {D_CHUNKS_A, D_CHUNKS_LT, D_CHUNKS_GL};
    const int hfs[3] = {D_HALFS_A, D_HALFS_LT, D_HALFS_GL};

    // ---- one device arena, every section 16-byte aligned
    auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
    size_t pos = 0;
    size_t o_prog[3];
    for (int k = 0; k < 3; ++k) { o_prog[k] = pos; pos += al((size_t)(h[hfs[k]] / 2 + 4) * 4); }
    const size_t o_next = pos; pos += al(next_lit.size() * 2);
    const size_t o_legal = pos; pos += al(legal_lit.size() * 2);
    const size_t o_glit = pos; pos += al(goal_lit.size() * 2);
    const size_t o_gval = pos; pos += al(goal_val.size());
    const size_t o_init = pos; pos += al((size_t)P.state_words * 4);
    // wide games: the same literal tables with 32-bit entries
    std::vector<uint32_t> next32, legal32, goal32;
    size_t o_next32 = 0, o_legal32 = 0, o_glit32 = 0;
    if (P.wide) {
        next32.assign((size_t)P.nbw * 32, 1u);
        for (int b = 0; b < P.n_bases; ++b) next32[b] = (uint32_t)h[h[D_OFF_NEXT_LIT] + b];
        legal32.assign((size_t)P.lw_used * 32, 1u);
        int src = h[D_OFF_LEGAL_LIT];
        for (int r = 0; r < R; ++r)
            for (int a = 0; a < role_sizes[r]; ++a) legal32[(size_t)lwo[r] * 32 + a] = (uint32_t)h[src++];
        goal32.assign(std::max(n_goal, 1), 1u);
        for (int k = 0; k < n_goal; ++k) goal32[k] = (uint32_t)h[h[D_OFF_GOAL_LIT] + k];
        o_next32 = pos; pos += al(next32.size() * 4);
        o_legal32 = pos; pos += al(legal32.size() * 4);
        o_glit32 = pos; pos += al(goal32.size() * 4);
    }
    std::vector<uint8_t> arena(pos, 0);
    for (int k = 0; k < 3; ++k) {
        P.chunks[k] = h[chs[k]];
        P.words[k] = h[hfs[k]] / 2;
        memcpy(arena.data() + o_prog[k], h + h[offs[k]], (size_t)P.words[k] * 4);
    }
    memcpy(arena.data() + o_next, next_lit.data(), next_lit.size() * 2);
    memcpy(arena.data() + o_legal, legal_lit.data(), legal_lit.size() * 2);
    memcpy(arena.data() + o_glit, goal_lit.data(), goal_lit.size() * 2);
    memcpy(arena.data() + o_gval, goal_val.data(), goal_val.size());
    memcpy(arena.data() + o_init, h + h[D_OFF_INIT], (size_t)P.state_words * 4);
    if (P.wide) {
        memcpy(arena.data() + o_next32, next32.data(), next32.size() * 4);
        memcpy(arena.data() + o_legal32, legal32.data(), legal32.size() * 4);
        memcpy(arena.data() + o_glit32, goal32.data(), goal32.size() * 4);
    }
    cudaError_t e = cudaMalloc(&pn->d_arena, pos);
    if (e != cudaSuccess) { gz_set_error("cudaMalloc: %s", cudaGetErrorString(e)); delete pn; return GZ_ERR_CUDA; }
    e = cudaMemcpy(pn->d_arena, arena.data(), pos, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { gz_set_error("cudaMemcpy: %s", cudaGetErrorString(e)); cudaFree(pn->d_arena); delete pn; return GZ_ERR_CUDA; }
    const uint8_t* base = (const uint8_t*)pn->d_arena;
    for (int k = 0; k < 3; ++k) P.prog[k] = (const uint32_t*)(base + o_prog[k]);
    P.next_lit = (const uint16_t*)(base + o_next);
    P.legal_lit = (const uint16_t*)(base + o_legal);
    P.goal_lit = (const uint16_t*)(base + o_glit);
    P.goal_val = (const uint8_t*)(base + o_gval);
    P.init_words = (const uint32_t*)(base + o_init);
    P.next_lit32 = P.wide ? (const uint32_t*)(base + o_next32) : nullptr;
    P.legal_lit32 = P.wide ? (const uint32_t*)(base + o_legal32) : nullptr;
    P.goal_lit32 = P.wide ? (const uint32_t*)(base + o_glit32) : nullptr;
    int rc = configure_launch(pn, h);
    if (rc == GZ_OK) rc = build_group_tables(pn, h);
    if (rc != GZ_OK) { cudaFree(pn->d_arena); delete pn; return rc; }
    pn->spec = gz_spec_find(fnv1a64(h, (size_t)n_words * 4));
    *out = pn;
    return GZ_OK;
}

extern "C" void gz_propnet_destroy(gz_propnet* pn) {
    if (!pn) return;
    cudaSetDevice(pn->device);
    if (pn->d_arena) cudaFree(pn->d_arena);
    if (pn->d_scratch) cudaFree(pn->d_scratch);
    for (auto& ws : pn->wide_scratch) cudaFree(ws.second);
    for (auto& ws : pn->group_scratch) cudaFree(ws.second);
    for (auto& gs : pn->goal_scratch) cudaFree(gs.mem);
    if (pn->d_group_tables) cudaFree(pn->d_group_tables);
    if (pn->host_stream) cudaStreamDestroy(pn->host_stream);
    delete pn;
}

extern "C" int gz_propnet_set_spec_min_batch(gz_propnet* pn, int64_t min_batch) {
    if (!pn) { gz_set_error("null propnet"); return GZ_ERR_ARG; }
    pn->spec_min_batch = min_batch;
    return GZ_OK;
}

// Batches of at least `min_batch` states (default 1024 = one group) go through the group kernel; < 0 disables it.
extern "C" int gz_propnet_set_group_min_batch(gz_propnet* pn, int64_t min_batch) {
    if (!pn) { gz_set_error("null propnet"); return GZ_ERR_ARG; }
    pn->group_min_batch = min_batch;
    return GZ_OK;
}

extern "C" const uint32_t* gz_propnet_init_words(const gz_propnet* pn) { return pn ? pn->P.init_words : nullptr; }

extern "C" int64_t gz_propnet_info(const gz_propnet* pn, int what) {
    if (!pn) return -1;
    const PnParams& P = pn->P;
    switch (what) {
        case GZ_INFO_N_BASES: return P.n_bases;
        case GZ_INFO_N_ROLES: return P.n_roles;
        case GZ_INFO_STATE_WORDS: return P.state_words;
        case GZ_INFO_LEGAL_WORDS: return P.legal_words;
        case GZ_INFO_VAL_WORDS: return P.val_words;
        case GZ_INFO_WARPS_PER_CTA: return pn->warps_per_cta;
        case GZ_INFO_SMEM_BYTES: return (int64_t)pn->smem_bytes;
        case GZ_INFO_GRID: return pn->grid_max;
        case GZ_INFO_STAGED_MASK: return P.staged[0] | (P.staged[1] << 1) | (P.staged[2] << 2);
        case GZ_INFO_N_INPUTS: return P.n_inputs;
        case GZ_INFO_SPECIALISED: return pn->spec ? 1 : 0;
        default: break;
    }
    if (what >= GZ_INFO_ROLE_SIZE0 && what < GZ_INFO_ROLE_SIZE0 + P.n_roles) return P.role_sizes[what - GZ_INFO_ROLE_SIZE0];
    if (what >= GZ_INFO_LEGAL_WORD_OFF0 && what <= GZ_INFO_LEGAL_WORD_OFF0 + P.n_roles) return P.legal_word_off[what - GZ_INFO_LEGAL_WORD_OFF0];
    return -1;
}

static bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

template <int MODE>
static int launch(gz_propnet* pn, int64_t n, int64_t slot0, const uint32_t* states_in, const int16_t* moves_in,
                  uint32_t* states_out, uint32_t* legal, uint8_t* terminal, uint8_t* goals,
                  int16_t* moves_out, unsigned long long* transitions, uint64_t seed, uint64_t step,
                  int goals_mode, void* stream, uint32_t* own_scratch = nullptr, const int* n_dev = nullptr) {
    if (!pn) { gz_set_error("null propnet"); return GZ_ERR_ARG; }
    if (n < 0) { gz_set_error("negative batch"); return GZ_ERR_ARG; }
    if (n == 0) return GZ_OK;
    if (!states_in || !legal || !terminal || !goals || (MODE != MODE_EVAL && !states_out) ||
        (MODE == MODE_TRANSITION && !moves_in)) {
        gz_set_error("null device pointer");
        return GZ_ERR_ARG;
    }
    if (!aligned16(states_in) || !aligned16(legal) || (states_out && !aligned16(states_out))) {
        gz_set_error("state / legal arrays must be 16-byte aligned");
        return GZ_ERR_ARG;
    }
    if (goals_mode != GZ_GOALS_ALWAYS && goals_mode != GZ_GOALS_TERMINAL_ONLY && goals_mode != GZ_GOALS_NEVER_INTERNAL) {
        gz_set_error("bad goals_mode");
        return GZ_ERR_ARG;
    }
    GZ_CUDA(cudaSetDevice(pn->device));
    const int64_t n_tiles = (n + 31) / 32;
    cudaStream_t st = (cudaStream_t)stream;
    if (pn->spec && pn->spec_min_batch >= 0 && n >= pn->spec_min_batch && !n_dev && goals_mode != GZ_GOALS_NEVER_INTERNAL) {
        // large batch of a game with a specialised kernel: one thread per 32-game tile
        SpecArgs a;
        a.n = n; a.slot0 = slot0; a.states_in = states_in; a.moves_in = moves_in; a.states_out = states_out;
        a.legal = legal; a.terminal = terminal; a.goals = goals; a.moves_out = moves_out;
        a.transitions = transitions; a.seed = seed; a.step = step; a.goals_mode = goals_mode;
        const int64_t n_batches = (n + 1023) / 1024;
        const int grid = (int)std::min<int64_t>((n_batches + SPEC_WARPS - 1) / SPEC_WARPS, pn->n_sms);
        if (pn->spec->launch(MODE, a, grid, st) != 0) {
            gz_set_error("specialised kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            return GZ_ERR_CUDA;
        }
        return GZ_OK;
    }
    if (pn->group_min_batch >= 0 && n >= pn->group_min_batch && !n_dev) {
        // a thousand games and more: the group kernel (one tile per lane, chunks of a level over a CTA / cluster)
        static std::mutex gmu;
        const int64_t n_groups = (n + GROUP_GAMES - 1) / GROUP_GAMES;
        const int clusters = (int)std::min<int64_t>(n_groups, pn->group_clusters);
        uint32_t* scratch = own_scratch;
        if (!scratch) {
            std::lock_guard<std::mutex> lk(gmu);
            for (auto& ws : pn->group_scratch) if (ws.first == st) scratch = ws.second;
            if (!scratch) {
                const size_t bytes = (size_t)pn->group_clusters * (size_t)pn->group_stride * 4;
                cudaError_t e = cudaMalloc((void**)&scratch, bytes);
                if (e != cudaSuccess) { gz_set_error("cudaMalloc of %.1f MB propnet group scratch failed: %s", bytes / 1e6, cudaGetErrorString(e)); return GZ_ERR_CUDA; }
                pn->group_scratch.emplace_back(st, scratch);
            }
        }
        // few groups (the PUCT step of a wide game: 1024 games = ONE group): a cluster of 16 CTAs (non-portable
        // size) spreads a level's chunks over twice as many SMs; large batches keep 8, which divides 148 SMs better
        int cl = pn->group_cl;
        if (cl == 8 && pn->P.wide && pn->group_cl16_ok && n_groups <= (int64_t)(pn->n_sms / 16)) cl = 16;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(clusters * cl));
        cfg.blockDim = dim3(GROUP_WARPS * 32);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cl;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e;
        const int64_t stride = pn->group_stride;
        const int* nd = nullptr;
#define GZ_GROUP_LAUNCH(W, C)                                                                                          \
        e = cudaLaunchKernelEx(&cfg, pn_group_kernel<MODE, W, C>, pn->P, pn->GT, n, slot0, states_in, moves_in, states_out, \
                               legal, terminal, goals, moves_out, transitions, seed, step, goals_mode, scratch, stride, nd)
        if (pn->P.wide) { if (cl == 16) GZ_GROUP_LAUNCH(true, 16); else if (cl == 8) GZ_GROUP_LAUNCH(true, 8); else GZ_GROUP_LAUNCH(true, 1); }
        else { if (pn->group_cl == 8) GZ_GROUP_LAUNCH(false, 8); else GZ_GROUP_LAUNCH(false, 1); }
#undef GZ_GROUP_LAUNCH
        if (e != cudaSuccess) { gz_set_error("group kernel launch failed: %s", cudaGetErrorString(e)); return GZ_ERR_CUDA; }
        return GZ_OK;
    }
    if (pn->P.wide) {
        static std::mutex mu;
        uint32_t* scratch = own_scratch;
        if (!scratch) {
            std::lock_guard<std::mutex> lk(mu);
            for (auto& ws : pn->wide_scratch) if (ws.first == st) scratch = ws.second;
            if (!scratch) {
                const size_t bytes = (size_t)pn->wide_grid * (size_t)pn->wide_stride * 4;
                cudaError_t e = cudaMalloc((void**)&scratch, bytes);
                if (e != cudaSuccess) { gz_set_error("cudaMalloc of %.1f MB propnet scratch failed: %s", bytes / 1e6, cudaGetErrorString(e)); return GZ_ERR_CUDA; }
                GZ_CUDA(cudaMemset(scratch, 0, bytes));
                pn->wide_scratch.emplace_back(st, scratch);
            }
        }
        const int grid = (int)std::min<int64_t>(n_tiles, pn->wide_grid);
        pn_wide_kernel<MODE><<<grid, WIDE_WARPS * 32, 0, st>>>(pn->P, n, slot0, states_in, moves_in, states_out, legal,
                                                            terminal, goals, moves_out, transitions, seed, step,
                                                            goals_mode, scratch, pn->wide_stride, n_dev);
        GZ_CUDA(cudaGetLastError());
        return GZ_OK;
    }
    // Small batches (the PUCT pipeline: a few thousand games) do not fill the machine with one warp per
    // tile and are bound by the serial latency of the gate programs: give every tile a CTA of 4 or 8
    // warps that share its value array. Large batches keep one warp per tile, many tiles per CTA.
    int wpt = 1;
    if (n_tiles <= (int64_t)pn->grid_max) wpt = 8;
    else if (n_tiles <= (int64_t)pn->grid_max * 2) wpt = 4;
    if (wpt > 1) {
        const size_t smem = (size_t)pn->P.staged_words * 4 + (size_t)pn->P.warp_words * 4;
        const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)pn->grid_max * 8);
        if (wpt == 8)
            pn_kernel<MODE, 8><<<grid, 8 * 32, smem, st>>>(pn->P, pn->GT, n, slot0, states_in, moves_in, states_out, legal,
                                                          terminal, goals, moves_out, transitions, seed, step, goals_mode, n_dev);
        else
            pn_kernel<MODE, 4><<<grid, 4 * 32, smem, st>>>(pn->P, pn->GT, n, slot0, states_in, moves_in, states_out, legal,
                                                          terminal, goals, moves_out, transitions, seed, step, goals_mode, n_dev);
    } else {
        // one persistent CTA per SM; medium batches shrink the CTA so that tiles spread over all SMs
        const int block_warps = (int)std::max<int64_t>(
            1, std::min<int64_t>(pn->warps_per_cta, (n_tiles + pn->grid_max - 1) / pn->grid_max));
        const int grid2 = (int)std::min<int64_t>((n_tiles + block_warps - 1) / block_warps, pn->grid_max);
        const size_t smem = (size_t)pn->P.staged_words * 4 + (size_t)block_warps * pn->P.warp_words * 4;
        pn_kernel<MODE, 1><<<grid2, block_warps * 32, smem, st>>>(
            pn->P, pn->GT, n, slot0, states_in, moves_in, states_out, legal, terminal, goals, moves_out, transitions,
            seed, step, goals_mode, n_dev);
    }
    GZ_CUDA(cudaGetLastError());
    return GZ_OK;
}

static bool use_deferred_goals(const gz_propnet* pn, int64_t n, int goals_mode);
static int deferred_goals(gz_propnet* pn, int64_t n, const uint32_t* states, const uint8_t* terminal, uint8_t* goals,
                          void* stream);

extern "C" int gz_sm_eval(gz_propnet* pn, int64_t n, const uint32_t* d_states, uint32_t* d_legal,
                          uint8_t* d_terminal, uint8_t* d_goals, int goals_mode, void* stream) {
    return launch<MODE_EVAL>(pn, n, 0, d_states, nullptr, nullptr, d_legal, d_terminal, d_goals, nullptr,
                             nullptr, 0, 0, goals_mode, stream);
}

extern "C" int gz_sm_transition(gz_propnet* pn, int64_t n, const uint32_t* d_states, const int16_t* d_moves,
                                uint32_t* d_next, uint32_t* d_legal, uint8_t* d_terminal, uint8_t* d_goals,
                                int goals_mode, void* stream) {
    if (pn && use_deferred_goals(pn, n, goals_mode)) {
        int rc = launch<MODE_TRANSITION>(pn, n, 0, d_states, d_moves, d_next, d_legal, d_terminal, d_goals,
                                         nullptr, nullptr, 0, 0, GZ_GOALS_NEVER_INTERNAL, stream);
        if (rc != GZ_OK) return rc;
        return deferred_goals(pn, n, d_next, d_terminal, d_goals, stream);
    }
    return launch<MODE_TRANSITION>(pn, n, 0, d_states, d_moves, d_next, d_legal, d_terminal, d_goals,
                                   nullptr, nullptr, 0, 0, goals_mode, stream);
}

// ---- deferred goal pass ---------------------------------------------------------------------------
// The reference reads goal values only of terminal states (src/cpp/puct/node.cpp:153-221; its random playouts,
// src/cpp/pyobjects/gdltransformer_impl.cpp:50-124, never do). A bit-sliced kernel evaluates a cone for the whole
// 32-game tile as soon as ONE game of the tile needs it, and reversi's goal cone (piece counting: 596 of its 795
// chunks, too large to stage in shared memory) then costs more than everything else together: with games at all
// plies at once 40 % of the tiles hold a terminal state although 1.7 % of the games do. For goals_mode
// GZ_GOALS_TERMINAL_ONLY and a game with a deep goal cone the main pass therefore skips the cone, and the terminal
// states of the batch are compacted, evaluated as dense tiles and their goal values scattered back — the same
// values, 20x less goal work (the PUCT engine does the same inside its step, gz_puct.cu).
__global__ void dg_compact_kernel(const uint8_t* __restrict__ terminal, int64_t n, int* __restrict__ idx,
                                  int* __restrict__ count) {
    // unordered compaction: one atomic per warp
    const int lane = threadIdx.x & 31;
    for (int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~(int64_t)31; i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = i0 + lane;
        const bool t = i < n && terminal[i] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, t);
        if (!m) continue;
        int base = 0;
        if (lane == 0) base = atomicAdd(count, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (t) idx[base + __popc(m & ((1u << lane) - 1u))] = (int)i;
    }
}
__global__ void dg_gather_kernel(const uint32_t* __restrict__ states, const int* __restrict__ idx,
                                 const int* __restrict__ count, int SW, uint32_t* __restrict__ out) {
    const int64_t total = (int64_t)*count * SW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / SW;
        out[i] = states[(int64_t)idx[r] * SW + (i - r * SW)];
    }
}
__global__ void dg_scatter_kernel(const uint8_t* __restrict__ t_goals, const int* __restrict__ idx,
                                  const int* __restrict__ count, int R, uint8_t* __restrict__ goals) {
    const int64_t total = (int64_t)*count * R;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / R;
        goals[(int64_t)idx[r] * R + (i - r * R)] = t_goals[i];
    }
}

static bool use_deferred_goals(const gz_propnet* pn, int64_t n, int goals_mode) {
    return goals_mode == GZ_GOALS_TERMINAL_ONLY && pn->P.chunks[2] >= 64 && pn->deferred_goals_min_batch >= 0 &&
           n >= pn->deferred_goals_min_batch && !(pn->spec && pn->spec_min_batch >= 0 && n >= pn->spec_min_batch);
}

// goal values of the terminal states among `states` (flags in `terminal`), written into `goals`
static int deferred_goals(gz_propnet* pn, int64_t n, const uint32_t* states, const uint8_t* terminal, uint8_t* goals,
                          void* stream) {
    const PnParams& P = pn->P;
    cudaStream_t st = (cudaStream_t)stream;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_idx = al((size_t)n * 4), b_cnt = 256, b_states = al((size_t)n * P.state_words * 4),
                 b_legal = al((size_t)n * P.legal_words * 4), b_term = al((size_t)n), b_goals = al((size_t)n * P.n_roles);
    uint8_t* mem = nullptr;
    {
        static std::mutex mu;
        std::lock_guard<std::mutex> lk(mu);
        for (auto& gs : pn->goal_scratch)
            if (gs.st == st && gs.cap >= n) mem = (uint8_t*)gs.mem;
        if (!mem) {
            const size_t bytes = b_idx + b_cnt + b_states + b_legal + b_term + b_goals;
            cudaError_t e = cudaMalloc((void**)&mem, bytes);
            if (e != cudaSuccess) { gz_set_error("cudaMalloc of %.1f MB for the deferred goal pass failed: %s", bytes / 1e6, cudaGetErrorString(e)); return GZ_ERR_CUDA; }
            pn->goal_scratch.push_back({st, n, mem});
        }
    }
    int* d_idx = (int*)mem;
    int* d_count = (int*)(mem + b_idx);
    uint32_t* t_states = (uint32_t*)(mem + b_idx + b_cnt);
    uint32_t* t_legal = (uint32_t*)(mem + b_idx + b_cnt + b_states);
    uint8_t* t_term = mem + b_idx + b_cnt + b_states + b_legal;
    uint8_t* t_goals = t_term + b_term;
    GZ_CUDA(cudaMemsetAsync(d_count, 0, 4, st));
    const int cgrid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)pn->n_sms * 8);
    dg_compact_kernel<<<cgrid, 256, 0, st>>>(terminal, n, d_idx, d_count);
    dg_gather_kernel<<<pn->n_sms * 4, 256, 0, st>>>(states, d_idx, d_count, P.state_words, t_states);
    GZ_CUDA(cudaGetLastError());
    int rc = launch<MODE_EVAL>(pn, n, 0, t_states, nullptr, nullptr, t_legal, t_term, t_goals, nullptr, nullptr, 0, 0,
                               GZ_GOALS_ALWAYS, stream, nullptr, d_count);
    if (rc != GZ_OK) return rc;
    dg_scatter_kernel<<<pn->n_sms * 2, 256, 0, st>>>(t_goals, d_idx, d_count, P.n_roles, goals);
    GZ_CUDA(cudaGetLastError());
    return GZ_OK;
}

// Batches of at least `min_batch` states of a game with a deep goal cone evaluate it in a compacted second pass when
// goals are wanted for terminal states only (default 4096; < 0: never).
extern "C" int gz_propnet_set_deferred_goals_min_batch(gz_propnet* pn, int64_t min_batch) {
    if (!pn) { gz_set_error("null propnet"); return GZ_ERR_ARG; }
    pn->deferred_goals_min_batch = min_batch;
    return GZ_OK;
}

// Internal (not part of the C-ABI): wide games evaluated with a caller-owned scratch region of
// gz_propnet_wide_scratch_bytes() bytes, so that an engine's steps can be captured in a CUDA graph and
// engines on different streams never share a value array.
extern "C" int64_t gz_propnet_wide_scratch_bytes(const gz_propnet* pn) {
    if (!pn) return 0;
    // one caller-owned region serves whichever kernel a launch picks: the per-CTA value arrays of the wide
    // kernel and the per-cluster value arrays of the group kernel
    const int64_t wide = pn->P.wide ? (int64_t)pn->wide_grid * pn->wide_stride * 4 : 0;
    const int64_t group = pn->group_min_batch >= 0 ? (int64_t)pn->group_clusters * pn->group_stride * 4 : 0;
    return std::max(wide, group);
}
extern "C" int gz_sm_transition_scratch(gz_propnet* pn, int64_t n, const uint32_t* d_states, const int16_t* d_moves,
                                        uint32_t* d_next, uint32_t* d_legal, uint8_t* d_terminal, uint8_t* d_goals,
                                        int goals_mode, void* stream, uint32_t* d_scratch) {
    return launch<MODE_TRANSITION>(pn, n, 0, d_states, d_moves, d_next, d_legal, d_terminal, d_goals,
                                   nullptr, nullptr, 0, 0, goals_mode, stream, d_scratch);
}

// Internal: evaluate the first *d_n (<= n_max) states; the count stays on the device.
extern "C" int gz_sm_eval_devcount(gz_propnet* pn, int64_t n_max, const int* d_n, const uint32_t* d_states,
                                   uint32_t* d_legal, uint8_t* d_terminal, uint8_t* d_goals, int goals_mode,
                                   void* stream, uint32_t* d_scratch) {
    return launch<MODE_EVAL>(pn, n_max, 0, d_states, nullptr, nullptr, d_legal, d_terminal, d_goals, nullptr,
                             nullptr, 0, 0, goals_mode, stream, d_scratch, d_n);
}
extern "C" int gz_propnet_goal_chunks(const gz_propnet* pn) { return pn ? pn->P.chunks[2] : 0; }

extern "C" int gz_sm_playout_step(gz_propnet* pn, int64_t slot0, int64_t n, uint64_t seed, uint64_t step,
                                  uint32_t* d_states, uint32_t* d_legal, uint8_t* d_terminal,
                                  uint8_t* d_goals, int16_t* d_moves_out,
                                  unsigned long long* d_transitions, int goals_mode, void* stream) {
    if (pn && use_deferred_goals(pn, n, goals_mode)) {
        int rc = launch<MODE_PLAYOUT>(pn, n, slot0, d_states, nullptr, d_states, d_legal, d_terminal, d_goals,
                                      d_moves_out, d_transitions, seed, step, GZ_GOALS_NEVER_INTERNAL, stream);
        if (rc != GZ_OK) return rc;
        return deferred_goals(pn, n, d_states, d_terminal, d_goals, stream);
    }
    return launch<MODE_PLAYOUT>(pn, n, slot0, d_states, nullptr, d_states, d_legal, d_terminal, d_goals,
                                d_moves_out, d_transitions, seed, step, goals_mode, stream);
}

extern "C" int gz_sm_transition_host(gz_propnet* pn, int64_t n, const uint32_t* h_states,
                                     const int16_t* h_moves, uint32_t* h_next, uint32_t* h_legal,
                                     uint8_t* h_terminal, uint8_t* h_goals, int goals_mode) {
    if (!pn) { gz_set_error("null propnet"); return GZ_ERR_ARG; }
    if (n <= 0) return n == 0 ? GZ_OK : GZ_ERR_ARG;
    if (!h_states || !h_moves || !h_next || !h_legal || !h_terminal || !h_goals) {
        gz_set_error("null host pointer");
        return GZ_ERR_ARG;
    }
    GZ_CUDA(cudaSetDevice(pn->device));
    const PnParams& P = pn->P;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_states = al((size_t)n * P.state_words * 4);
    const size_t b_moves = al((size_t)n * P.n_roles * 2);
    const size_t b_legal = al((size_t)n * P.legal_words * 4);
    const size_t b_term = al((size_t)n);
    const size_t b_goals = al((size_t)n * P.n_roles);
    const size_t need = 2 * b_states + b_moves + b_legal + b_term + b_goals;
    if (need > pn->scratch_bytes) {
        if (pn->d_scratch) cudaFree(pn->d_scratch);
        pn->d_scratch = nullptr;
        pn->scratch_bytes = 0;
        GZ_CUDA(cudaMalloc(&pn->d_scratch, need));
        pn->scratch_bytes = need;
    }
    if (!pn->host_stream) GZ_CUDA(cudaStreamCreateWithFlags(&pn->host_stream, cudaStreamNonBlocking));
    uint8_t* p = (uint8_t*)pn->d_scratch;
    uint32_t* d_states = (uint32_t*)p; p += b_states;
    uint32_t* d_next = (uint32_t*)p; p += b_states;
    int16_t* d_moves = (int16_t*)p; p += b_moves;
    uint32_t* d_legal = (uint32_t*)p; p += b_legal;
    uint8_t* d_term = p; p += b_term;
    uint8_t* d_goals = p;
    cudaStream_t s = pn->host_stream;
    GZ_CUDA(cudaMemcpyAsync(d_states, h_states, (size_t)n * P.state_words * 4, cudaMemcpyHostToDevice, s));
    GZ_CUDA(cudaMemcpyAsync(d_moves, h_moves, (size_t)n * P.n_roles * 2, cudaMemcpyHostToDevice, s));
    int rc = gz_sm_transition(pn, n, d_states, d_moves, d_next, d_legal, d_term, d_goals, goals_mode, s);
    if (rc != GZ_OK) return rc;
    GZ_CUDA(cudaMemcpyAsync(h_next, d_next, (size_t)n * P.state_words * 4, cudaMemcpyDeviceToHost, s));
    GZ_CUDA(cudaMemcpyAsync(h_legal, d_legal, (size_t)n * P.legal_words * 4, cudaMemcpyDeviceToHost, s));
    GZ_CUDA(cudaMemcpyAsync(h_terminal, d_term, (size_t)n, cudaMemcpyDeviceToHost, s));
    GZ_CUDA(cudaMemcpyAsync(h_goals, d_goals, (size_t)n * P.n_roles, cudaMemcpyDeviceToHost, s));
    GZ_CUDA(cudaStreamSynchronize(s));
    return GZ_OK;
}
