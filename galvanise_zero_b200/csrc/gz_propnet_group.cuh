// gz_propnet_group.cuh — the "group" propnet kernel: 1024 games at a time, one game TILE per lane.
//
// pn_kernel / pn_wide_kernel give a tile of 32 games to a warp (or CTA) and let lane i evaluate gate i of a
// 32-gate chunk: every literal is a 4-byte gather from the tile's value array, and a tile's levels run one
// after the other with only 32 games of work in each instruction. That is the right shape for a handful of
// states (player mode), but batches of a thousand games and more are better served the other way round:
//
//   * a GROUP is 32 tiles = 1024 games; the value array holds, per live slot, the 32 tile words side by side
//     (V[slot][tile], 128 bytes) in an L2-resident scratch region;
//   * lane t of a warp works on tile t: a warp takes a whole 32-gate chunk and evaluates its gates one after
//     the other, every literal being ONE coalesced 128-byte load (the literal itself is warp-uniform, read
//     once per chunk row and handed round by shuffle), with the 32 gates of the chunk as 32 independent
//     accumulators — 32 loads in flight per warp instead of a dependent gather per level;
//   * the chunks of a level are spread over all warps of a CTA (narrow levels: reversi, breakthrough) or of
//     a whole thread-block cluster (wide levels: chess_15d 300 chunks per level, hexLG13 1400), which meet at
//     a CTA / cluster barrier per level; chunk offsets and level boundaries come from a table built at load
//     time, so nobody walks the chunk stream.
//
// Same device programs (csrc/gz_devformat.h), same results as the other kernels and as oracle/propnet_ref.c
// (tests/test_gpu_propnet.py runs every game through every kernel).
#ifndef GZ_PROPNET_GROUP_CUH
#define GZ_PROPNET_GROUP_CUH

#include <cooperative_groups.h>

#define GROUP_WARPS 16
#define GROUP_TILES 32
#define GROUP_GAMES (GROUP_TILES * 32)

#ifndef GZ_GROUP_TABLES_DEFINED
struct GroupTables {
    const uint32_t* desc[3];     // per chunk: word offset into the program | literal rows << 28
    const int* level_first[3];   // per level: index of its first chunk; [n_levels] = n_chunks
    int n_levels[3];
};
#endif

template <int CL>
__device__ __forceinline__ void group_sync() {
    if (CL == 1) __syncthreads();
    else cooperative_groups::this_cluster().sync();
}

__device__ __forceinline__ uint32_t gv_lit(const uint32_t* V, uint32_t lit, int lane) {
    return __ldcg(V + (size_t)(lit >> 1) * GROUP_TILES + lane) ^ (0u - (lit & 1u));
}

// Evaluates program k on the group's value array. gw = this warp's index among the CW warps that share the group.
template <bool WIDE, int CL>
__device__ __forceinline__ void run_group_program(const uint32_t* __restrict__ prog, const uint32_t* __restrict__ desc,
                                                  const int* __restrict__ level_first, int n_levels, uint32_t* V,
                                                  int lane, int gw, int CW) {
    for (int l = 0; l < n_levels; ++l) {
        const int c_end = __ldg(level_first + l + 1);
        for (int c = __ldg(level_first + l) + gw; c < c_end; c += CW) {
            const uint32_t d = __ldg(desc + c);
            const uint32_t* p = prog + (d & 0x0fffffffu);
            const int nl = (int)(d >> 28);
            uint32_t acc[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = 0xffffffffu;
            uint32_t out;
            // all literal rows of the chunk are fetched before any value (one memory round trip for the chunk's
            // program instead of one per row); rows beyond nl are the TRUE literal
            uint32_t rows[8];
            if (WIDE) {
                out = __ldg(p + 1 + lane);
#pragma unroll
                for (int k = 0; k < 8; ++k) rows[k] = k < nl ? __ldg(p + 33 + k * 32 + lane) : 0u;
                // the 32 values of a literal row are fetched as one batch (32 independent coalesced loads in flight)
                // and combined afterwards; interleaved load / and pairs kept only two or three loads in flight
                // (chess_15d 13.7 -> 22.0 M transitions/s)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (k < nl) {
                        uint32_t v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const uint32_t lit = __shfl_sync(0xffffffffu, rows[k], i);
                            v[i] = __ldcg(V + (size_t)(lit >> 1) * GROUP_TILES + lane);
                        }
                        const uint32_t inv = __ballot_sync(0xffffffffu, (rows[k] & 1u) != 0u);   // bit i: literal i is negated
#pragma unroll
                        for (int i = 0; i < 32; ++i) acc[i] &= v[i] ^ (0u - ((inv >> i) & 1u));
                    }
                }
            } else {
                const uint32_t ow = __ldg(p + 1 + (lane >> 1));
                out = (lane & 1) ? (ow >> 16) : (ow & 0xffffu);
#pragma unroll
                for (int k = 0; k < 4; ++k) rows[k] = k < (nl >> 1) ? __ldg(p + 17 + k * 32 + lane) : 0u;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k < (nl >> 1)) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const uint32_t w = __shfl_sync(0xffffffffu, rows[k], i);
                            acc[i] &= gv_lit(V, w & 0xffffu, lane) & gv_lit(V, w >> 16, lane);
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 32; ++i)
                __stcg(V + (size_t)__shfl_sync(0xffffffffu, out, i) * GROUP_TILES + lane, acc[i]);
        }
        group_sync<CL>();
    }
}

template <int MODE, bool WIDE, int CL>
__global__ void __launch_bounds__(GROUP_WARPS * 32)
pn_group_kernel(const __grid_constant__ PnParams P, const __grid_constant__ GroupTables T, int64_t n, int64_t slot0,
                const uint32_t* states_in, const int16_t* __restrict__ moves_in, uint32_t* states_out, uint32_t* legal,
                uint8_t* terminal, uint8_t* __restrict__ goals, int16_t* __restrict__ moves_out,
                unsigned long long* __restrict__ transitions, uint64_t seed, uint64_t step, int goals_mode,
                uint32_t* scratch, int64_t scratch_stride, const int* __restrict__ n_dev) {
    if (n_dev) n = (int64_t)*n_dev;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rank = CL == 1 ? 0 : (int)cooperative_groups::this_cluster().block_rank();
    const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
    const int CW = CL * GROUP_WARPS, gw = rank * GROUP_WARPS + warp;
    const int R = P.n_roles, SW = P.state_words, LW = P.legal_words;
    uint32_t* V = scratch + (int64_t)cluster_id * scratch_stride;       // [val_words][32]
    uint32_t* S = V + (size_t)P.val_words * GROUP_TILES;                 // [nbw * 32][32] next-state slices
    // [2][32] reset / pass masks per tile, behind the slices region (which also lends GROUP_TILES x roles x 32 words
    // to the moves of a warp's second tile, so it is at least that large: build_group_tables)
    const size_t slices_words = (size_t)P.nbw * 32 * GROUP_TILES > (size_t)GROUP_TILES * GZ_PN_MAX_ROLES * 32
                                    ? (size_t)P.nbw * 32 * GROUP_TILES : (size_t)GROUP_TILES * GZ_PN_MAX_ROLES * 32;
    uint32_t* M = S + slices_words;
    const int64_t n_groups = (n + GROUP_GAMES - 1) / GROUP_GAMES;
    for (int64_t group = cluster_id; group < n_groups; group += n_clusters) {
        const int64_t gbase = group * GROUP_GAMES;
        // ---- constants, input slots, moves
        if (gw == 0) __stcg(V + lane, 0xffffffffu);                       // slot 0 = TRUE
        if (MODE != MODE_EVAL)
            for (int i = gw * 32 + lane; i < P.n_inputs * GROUP_TILES; i += CW * 32)
                __stcg(V + (size_t)(1 + P.n_bases) * GROUP_TILES + i, 0u);
        int mv[GZ_PN_MAX_ROLES];
#pragma unroll
        for (int r = 0; r < GZ_PN_MAX_ROLES; ++r) mv[r] = -1;
        int my_tile = -1;
        if (MODE != MODE_EVAL) {
            // tile t of the group belongs to warp t % CW (CW is 16 or 128: one or two tiles per warp, or none);
            // a warp with two tiles handles the second one after the barrier below
            for (int t = gw; t < GROUP_TILES; t += CW) {
                const int64_t g0 = gbase + (int64_t)t * 32;
                const int64_t game = g0 + lane;
                const bool valid = game < n;
                uint32_t reset_mask = 0, pass_mask = 0;
                int m[GZ_PN_MAX_ROLES];
                if (MODE == MODE_PLAYOUT) {
                    const bool term = valid ? (terminal[game] != 0) : false;
                    reset_mask = __ballot_sync(0xffffffffu, term);
                    for (int r = 0; r < GZ_PN_MAX_ROLES; ++r) {
                        m[r] = -1;
                        if (r < R && valid && !term) {
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
                            m[r] = idx;
                        }
                        if (r < R && valid && moves_out) moves_out[game * R + r] = (int16_t)m[r];
                    }
                } else {
                    for (int r = 0; r < GZ_PN_MAX_ROLES; ++r) m[r] = (r < R && valid) ? (int)moves_in[game * R + r] : -1;
                    pass_mask = __ballot_sync(0xffffffffu, valid && m[0] < 0);
                }
                if (lane == 0) { __stcg(M + t, reset_mask); __stcg(M + GROUP_TILES + t, pass_mask); }
                if (t == gw) {
                    my_tile = t;
#pragma unroll
                    for (int r = 0; r < GZ_PN_MAX_ROLES; ++r) mv[r] = m[r];
                } else {
                    // second tile of this warp: its moves wait in the slices region (free until the next-state step)
                    for (int r = 0; r < R; ++r) __stcg(S + ((size_t)t * GZ_PN_MAX_ROLES + r) * 32 + lane, (uint32_t)m[r]);
                }
            }
        }
        // ---- packed states -> base slots: item (tile, state word)
        for (int it = gw; it < GROUP_TILES * P.nbw; it += CW) {
            const int t = it / P.nbw, j = it - t * P.nbw;
            const int64_t game = gbase + (int64_t)t * 32 + lane;
            const uint32_t x = game < n ? states_in[game * SW + j] : 0u;
            const uint32_t w = warp_transpose32(x, lane);
            const int b = j * 32 + lane;
            if (b < P.n_bases) __stcg(V + (size_t)(1 + b) * GROUP_TILES + t, w);
        }
        group_sync<CL>();

        if (MODE != MODE_EVAL) {
            // ---- one-hot the chosen actions into the input slots
            for (int t = gw; t < GROUP_TILES; t += CW) {
                for (int r = 0; r < R; ++r) {
                    const int m = (t == my_tile) ? mv[r] : (int)__ldcg(S + ((size_t)t * GZ_PN_MAX_ROLES + r) * 32 + lane);
                    if (m >= 0) atomicOr(V + (size_t)(P.input_slot0[r] + m) * GROUP_TILES + t, 1u << lane);
                }
            }
            group_sync<CL>();
            // ---- next-state cone
            run_group_program<WIDE, CL>(P.prog[0], T.desc[0], T.level_first[0], T.n_levels[0], V, lane, gw, CW);
            for (int it = gw; it < GROUP_TILES * P.nbw; it += CW) {
                const int t = it / P.nbw, j = it - t * P.nbw;
                const int64_t game = gbase + (int64_t)t * 32 + lane;
                const bool valid = game < n;
                const uint32_t valid_mask = __ballot_sync(0xffffffffu, valid);
                const int b = j * 32 + lane;
                const uint32_t lit = WIDE ? __ldg(P.next_lit32 + b) : (uint32_t)__ldg(P.next_lit + b);
                uint32_t w = __ldcg(V + (size_t)(lit >> 1) * GROUP_TILES + t) ^ (0u - (lit & 1u));
                if (MODE == MODE_PLAYOUT) {
                    const uint32_t reset_mask = __ldcg(M + t);
                    const uint32_t ib = (P.init_words[j] >> lane) & 1u;
                    w = (w & ~reset_mask) | (ib ? reset_mask : 0u);
                } else if (b < P.n_bases) {
                    const uint32_t pass_mask = __ldcg(M + GROUP_TILES + t);
                    w = (w & ~pass_mask) | (__ldcg(V + (size_t)(1 + b) * GROUP_TILES + t) & pass_mask);
                }
                w &= valid_mask;
                __stcg(S + (size_t)b * GROUP_TILES + t, w);
                const uint32_t x = warp_transpose32(w, lane);
                if (valid) states_out[game * SW + j] = x;
            }
            for (int it = gw; it < GROUP_TILES * (SW - P.nbw); it += CW) {
                const int t = it / (SW - P.nbw), j = P.nbw + it - t * (SW - P.nbw);
                const int64_t game = gbase + (int64_t)t * 32 + lane;
                if (game < n) states_out[game * SW + j] = 0u;
            }
            group_sync<CL>();
            for (int i = gw * 32 + lane; i < P.n_bases * GROUP_TILES; i += CW * 32)
                __stcg(V + (size_t)GROUP_TILES + i, __ldcg(S + i));
            group_sync<CL>();
        }

        // ---- views of the (new) state
        run_group_program<WIDE, CL>(P.prog[1], T.desc[1], T.level_first[1], T.n_levels[1], V, lane, gw, CW);
        const int64_t my_game = gbase + (int64_t)lane * 32;            // lane = tile here
        uint32_t tw_mine = gv_lit(V, (uint32_t)P.terminal_lit, lane);
        {
            const int64_t left = n - my_game;
            const uint32_t vm = left >= 32 ? 0xffffffffu : (left <= 0 ? 0u : ((1u << (int)left) - 1u));
            tw_mine &= vm;
        }
        const bool need_goals = (goals_mode == GZ_GOALS_ALWAYS) ||
                                (goals_mode == GZ_GOALS_TERMINAL_ONLY && __any_sync(0xffffffffu, tw_mine != 0u));
        if (need_goals) run_group_program<WIDE, CL>(P.prog[2], T.desc[2], T.level_first[2], T.n_levels[2], V, lane, gw, CW);

        for (int it = gw; it < GROUP_TILES * LW; it += CW) {
            const int t = it / LW, j = it - t * LW;
            const int64_t game = gbase + (int64_t)t * 32 + lane;
            const bool valid = game < n;
            const uint32_t valid_mask = __ballot_sync(0xffffffffu, valid);
            uint32_t x = 0u;
            if (j < P.lw_used) {
                const uint32_t lit = WIDE ? __ldg(P.legal_lit32 + j * 32 + lane) : (uint32_t)__ldg(P.legal_lit + j * 32 + lane);
                const uint32_t w = (__ldcg(V + (size_t)(lit >> 1) * GROUP_TILES + t) ^ (0u - (lit & 1u))) & valid_mask;
                x = warp_transpose32(w, lane);
            }
            if (valid) legal[game * LW + j] = x;
        }
        for (int t = gw; t < GROUP_TILES; t += CW) {
            const int64_t game = gbase + (int64_t)t * 32 + lane;
            const bool valid = game < n;
            const uint32_t valid_mask = __ballot_sync(0xffffffffu, valid);
            const uint32_t tw = __shfl_sync(0xffffffffu, tw_mine, t);
            const bool is_term = (tw >> lane) & 1u;
            if (valid) terminal[game] = is_term ? 1 : 0;
            for (int r = 0; r < R; ++r) {
                uint32_t gv = 255u;
                if (need_goals) {
                    for (int k = P.goal_off[r]; k < P.goal_off[r + 1]; ++k) {
                        const uint32_t lit = WIDE ? __ldg(P.goal_lit32 + k) : (uint32_t)__ldg(P.goal_lit + k);
                        const uint32_t w = __ldcg(V + (size_t)(lit >> 1) * GROUP_TILES + t) ^ (0u - (lit & 1u));
                        if ((w >> lane) & 1u) gv = P.goal_val[k];
                    }
                }
                if (goals_mode == GZ_GOALS_TERMINAL_ONLY && !is_term) gv = 255u;
                if (valid) goals[game * R + r] = (uint8_t)gv;
            }
            if (MODE == MODE_PLAYOUT && transitions && lane == 0) {
                const int cnt = __popc(valid_mask & ~__ldcg(M + t));
                if (cnt) atomicAdd(transitions, (unsigned long long)cnt);
            }
        }
        group_sync<CL>();
    }
}

#endif  // GZ_PROPNET_GROUP_CUH
