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
template <int J>
__device__ __forceinline__ void spec_transpose_stage(uint32_t (&a)[32]) {
    constexpr uint32_t m = J == 4 ? 0x0F0F0F0Fu : J == 2 ? 0x33333333u : 0x55555555u;
#pragma unroll
    for (int base = 0; base < 32; base += 2 * J) {
#pragma unroll
        for (int k = base; k < base + J; ++k) {               // constant indices: a[] stays in registers
            const uint32_t x = a[k], y = a[k + J];
            if (J == 16) {
                a[k] = __byte_perm(x, y, 0x5410);
                a[k + J] = __byte_perm(x, y, 0x7632);
            } else if (J == 8) {
                a[k] = __byte_perm(x, y, 0x6240);
                a[k + J] = __byte_perm(x, y, 0x7351);
            } else {
                a[k] = (x & m) | ((y << J) & ~m);
                a[k + J] = ((x >> J) & m) | (y & ~m);
            }
        }
    }
}
__device__ __forceinline__ void spec_transpose32(uint32_t (&a)[32]) {
    spec_transpose_stage<16>(a);
    spec_transpose_stage<8>(a);
    spec_transpose_stage<4>(a);
    spec_transpose_stage<2>(a);
    spec_transpose_stage<1>(a);
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

// Staging. A lane's tile is 32 consecutive rows in HBM; a warp's 32 tiles are contiguous. Rows move
// between HBM and the lanes in GROUPS of SPEC_GR rows per tile: the warp copies, for every tile, the
// group's SPEC_GR * RW contiguous words with coalesced 16-byte cp.async / stores into a per-warp
// shared-memory block [32 tiles][SPEC_GR * RW + 4 pad], and every lane then reads / writes its own
// tile's rows there (128-bit accesses are bank-conflict free with the 4-word pad). Staging a quarter of
// the rows at a time keeps the block at 12.8 KB per warp, so that eight warps fit on an SM.
#define SPEC_GR 8
#define SPEC_NG (32 / SPEC_GR)

template <int RW>
__device__ __forceinline__ void spec_fill_group(uint32_t* buf, const uint32_t* __restrict__ base, int h,
                                                int64_t warp_rows, int lane) {
    // base: first word of the warp's first row; group h covers rows [h*GR, h*GR+GR) of each tile
    constexpr int GW = SPEC_GR * RW, CPT = GW / 4;             // words, 16-byte chunks per tile and group
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(buf);
#pragma unroll 4
    for (int c = lane; c < 32 * CPT; c += 32) {
        const int t = c / CPT, o = (c - t * CPT) * 4;
        const int row = t * 32 + h * SPEC_GR + o / RW;
        const uint32_t dst = sbase + (uint32_t)(t * (GW + 4) + o) * 4u;
        const uint32_t* src = base + (int64_t)(t * 32 + h * SPEC_GR) * RW + o;
        const int bytes = row < warp_rows ? 16 : 0;             // rows beyond the batch read as zero
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void spec_fill_wait() {
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncwarp();
}

template <int RW>
__device__ __forceinline__ void spec_drain_group(const uint32_t* buf, uint32_t* __restrict__ base, int h,
                                                 int64_t warp_rows, int lane) {
    constexpr int GW = SPEC_GR * RW, CPT = GW / 4;
    __syncwarp();
#pragma unroll 4
    for (int c = lane; c < 32 * CPT; c += 32) {
        const int t = c / CPT, o = (c - t * CPT) * 4;
        const int row = t * 32 + h * SPEC_GR + o / RW;
        if (row < warp_rows)
            *reinterpret_cast<uint4*>(base + (int64_t)(t * 32 + h * SPEC_GR) * RW + o) =
                *reinterpret_cast<const uint4*>(buf + t * (GW + 4) + o);
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

#define SPEC_WARPS 8

// G: the generated game description (csrc/gz_spec_<game>.cu), see tools/gen_spec_kernel.py.
//
// Code size matters here: the gate networks are straight-line by nature (a few thousand LOP3s that
// every warp runs once per batch), so everything repetitive — the per-game move chooser, the word-by-word
// transposes back to packed rows — is a real loop whose body stays in the instruction cache (the first,
// fully unrolled version spent 55 % of its cycles waiting for instruction fetch). Warps run unsynchronised
// so that the memory phases of one overlap the arithmetic of the others.
template <typename G, int MODE>
__global__ void __launch_bounds__(SPEC_WARPS * 32, 1) spec_kernel(const SpecArgs A) {
    extern __shared__ __align__(16) uint32_t spec_smem[];
    constexpr int SW = G::SW, LW = G::LW, R = G::R, NBW = G::NBW, LWU = G::LWU;
    constexpr int SGW = SPEC_GR * SW, LGW = SPEC_GR * LW;        // words per tile and group
    constexpr int BUFW = 32 * ((LGW > SGW ? LGW : SGW) + 4);     // words per warp buffer
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* buf = spec_smem + (size_t)warp * BUFW;
    uint32_t* mine_s = buf + lane * (SGW + 4);                   // this lane's rows of the group (state)
    uint32_t* mine_l = buf + lane * (LGW + 4);                   // ... (legal)

    const int64_t n = A.n;
    const int64_t n_batches = (n + 1023) >> 10;                  // a warp's unit: 32 tiles = 1024 games
    const int64_t per_round = (int64_t)gridDim.x * SPEC_WARPS;
    const int64_t rounds = (n_batches + per_round - 1) / per_round;
    for (int64_t round = 0; round < rounds; ++round) {
        const int64_t batch = round * per_round + (int64_t)blockIdx.x * SPEC_WARPS + warp;
        const bool active = batch < n_batches;                   // idle warps still meet the barriers
        const int64_t w0 = active ? batch << 10 : 0;             // first game of the warp
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

        if (MODE == SPEC_PLAYOUT) {
            // ---- one uniformly random legal action per role and game, from the stored legal rows
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
#pragma unroll 1
            for (int h = 0; h < SPEC_NG; ++h) {
                spec_fill_group<LW>(buf, A.legal + w0 * LW, h, warp_rows, lane);
                spec_fill_wait();
#pragma unroll 2
                for (int ii = 0; ii < SPEC_GR; ++ii) {
                    const int i = h * SPEC_GR + ii;
                    const bool term = (reset_mask >> i) & 1u;
                    const bool valid = i < rows;
                    uint32_t lw[LW];
#pragma unroll
                    for (int q = 0; q < LW / 4; ++q) {
                        const uint4 v = *reinterpret_cast<const uint4*>(mine_l + ii * LW + q * 4);
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
            }
        } else if (MODE == SPEC_TRANSITION) {
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

        // ---- packed state rows -> registers (row-major), then 32x32 transposes in place -> bit-sliced bases
        uint32_t b[NBW * 32];
        spec_for<0, SPEC_NG>([&](auto H) {
            constexpr int h = decltype(H)::value;
            spec_fill_group<SW>(buf, A.states_in + w0 * SW, h, warp_rows, lane);
            spec_fill_wait();
#pragma unroll
            for (int ii = 0; ii < SPEC_GR; ++ii) {
#pragma unroll
                for (int q = 0; q < (NBW + 3) / 4; ++q) {
                    const uint4 v = *reinterpret_cast<const uint4*>(mine_s + ii * SW + q * 4);
                    const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (q * 4 + e < NBW) b[(q * 4 + e) * 32 + h * SPEC_GR + ii] = vv[e];
                }
            }
            __syncwarp();
        });
        spec_for<0, NBW>([&](auto J) {
            constexpr int j = decltype(J)::value;
            uint32_t a[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = b[j * 32 + i];
            spec_transpose32(a);
#pragma unroll
            for (int i = 0; i < 32; ++i) b[j * 32 + i] = a[i];
        });

        uint32_t sl[(LWU > NBW ? LWU : NBW) * 32];              // bit-sliced outputs, indexed in loops (local memory)

        if (MODE != SPEC_EVAL) {
            // ---- next state: straight-line gate network per state word ...
            typename G::Decode dec;
            G::decode_moves(mv, dec);

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
            // ... the new bases replace the old ones ...
#pragma unroll
            for (int i = 0; i < NBW * 32; ++i) b[i] = sl[i];
            // ... and go back to packed rows: transpose word by word in place, then drain group by group
#pragma unroll 1
            for (int j = 0; j < NBW; ++j) {
                uint32_t o[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = sl[j * 32 + i];
                spec_transpose32(o);
#pragma unroll
                for (int i = 0; i < 32; ++i) sl[j * 32 + i] = o[i];
            }
#pragma unroll 1
            for (int h = 0; h < SPEC_NG; ++h) {
#pragma unroll
                for (int ii = 0; ii < SPEC_GR; ++ii) {
#pragma unroll
                    for (int q = 0; q < SW / 4; ++q) {
                        uint32_t vv[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            vv[e] = (q * 4 + e < NBW) ? sl[(q * 4 + e) * 32 + h * SPEC_GR + ii] : 0u;
                        *reinterpret_cast<uint4*>(mine_s + ii * SW + q * 4) = make_uint4(vv[0], vv[1], vv[2], vv[3]);
                    }
                }
                spec_drain_group<SW>(buf, A.states_out + w0 * SW, h, warp_rows, lane);
            }
        }

        // ---- views of the (new) state: legal masks, terminal, goals

        spec_for<0, LWU>([&](auto J) {
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
        for (int j = 0; j < LWU; ++j) {
            uint32_t o[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = sl[j * 32 + i];
            spec_transpose32(o);
#pragma unroll
            for (int i = 0; i < 32; ++i) sl[j * 32 + i] = o[i];
        }
#pragma unroll 1
        for (int h = 0; h < SPEC_NG; ++h) {
#pragma unroll
            for (int ii = 0; ii < SPEC_GR; ++ii) {
#pragma unroll
                for (int q = 0; q < LW / 4; ++q) {
                    uint32_t vv[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        vv[e] = (q * 4 + e < LWU) ? sl[(q * 4 + e) * 32 + h * SPEC_GR + ii] : 0u;
                    *reinterpret_cast<uint4*>(mine_l + ii * LW + q * 4) = make_uint4(vv[0], vv[1], vv[2], vv[3]);
                }
            }
            spec_drain_group<LW>(buf, A.legal + w0 * LW, h, warp_rows, lane);
        }

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
    constexpr int SGW = SPEC_GR * G::SW, LGW = SPEC_GR * G::LW;
    constexpr size_t smem = (size_t)SPEC_WARPS * 32 * ((LGW > SGW ? LGW : SGW) + 4) * 4;
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
