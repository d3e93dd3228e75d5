// gz_encoder_col.cuh — the column-owned plane encoder for large batches.
//
// encode_fast_kernel (gz_encoder.cu) interprets one descriptor per output ELEMENT per state: ~30 instructions per
// element (descriptor fetch and decode, three divergent kinds), which bounds it at 15-25 % of the HBM roofline
// although the work is a pure stream (profiles/r02_bench_propnet_games.jsonl). Here a thread OWNS one 8-element
// column of the plane row — the same 8 descriptors for every state — decodes them once into registers (word
// offset, shift, mask, multiplier: a bit of the state, a bit of the OR of the previous states, a control
// channel's value or a constant zero all become `((word >> shift) & mask) * mul`) and then streams the states of
// the CTA's tile through them: per 8 output elements 8 shared-memory loads, 24 ALU instructions and one 128-bit
// store (bf16; two for f32). Thread t writes flat chunk t + i * (threads in use), so a warp's stores are one
// contiguous 512-byte (bf16) run whatever the row size; a thread's column never changes because the stride is a
// multiple of the chunks per row.
//
// Rows that are not a multiple of 8 elements (hexLG13: 845, breakthroughSmall: 180) are handled in UNITS of
// U = 8 / gcd(8, row) consecutive states: a unit is a multiple of 8 elements, its descriptor pattern is the same for
// every unit, and a chunk that straddles two states simply reads its elements from two staged states (the state
// within the unit is part of the descriptor). States beyond the batch's last complete unit are written element-wise.
//
// Results are those of encode_kernel / encode_fast_kernel (tests/test_encoder.py runs all against the C oracle
// and the reference's compiled toChannels); tests/hostsim/ctasim_encoder.cpp compiles this header for the CPU.
#ifndef GZ_ENCODER_COL_CUH
#define GZ_ENCODER_COL_CUH

#include "gz_encoder.cuh"

#define ENCC_THREADS 256
// (the CPU validation build, tests/hostsim/ctasim_encoder.cpp, substitutes a pointer to its own buffer)
#ifndef ENCC_DYNAMIC_SMEM
#define ENCC_DYNAMIC_SMEM(name) extern __shared__ __align__(16) uint32_t name[]
#endif

// Shared-memory addresses of the staged words are kept as plain 32-bit shared-window addresses, so that the inner
// loop advances them with one add each and loads with ld.shared (the CPU validation build uses host pointers).
#ifndef ENCC_HOST_BUILD
typedef uint32_t encc_addr_t;
__device__ __forceinline__ encc_addr_t encc_addr(const uint32_t* p) { return (encc_addr_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t encc_lds(encc_addr_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
#else
typedef uintptr_t encc_addr_t;
static inline encc_addr_t encc_addr(const uint32_t* p) { return (encc_addr_t)p; }
static inline uint32_t encc_lds(encc_addr_t a) { return *reinterpret_cast<const uint32_t*>(a); }
#endif

// the row's element descriptors (shared with the description in gz_encoder.cu: kind in bits 30-31 — 0 zero, 1 bit
// of the state, 2 bit of the OR of the previous states, 3 control channel —, payload in bits 0-29)
// (units: bits 27-29 hold the state within the unit, the payload is 27 bits)
__device__ __forceinline__ void encc_build_desc(const EncParams& E, uint32_t* desc, int row_elems, int c_pad,
                                                bool has_prev, int tid, int nthreads, int U) {
    const int n_ch = E.total / E.channel_size;
    for (int e0 = tid; e0 < U * row_elems; e0 += nthreads) {
        const int su = e0 / row_elems, le0 = e0 - su * row_elems;
        int le = le0;
        uint32_t d = 0u;
        bool live = true;
        if (c_pad > 0) {                                  // [H*W][c_pad] rows: element = (cell, channel)
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
        if (d) d |= (uint32_t)su << 27;
        desc[e0] = d;
    }
}

template <typename OutT>
struct EnccColumn {          // one thread's column: 8 decoded descriptors (addr: the word of the thread's first state)
    encc_addr_t addr[8];
    uint32_t sh[8], msk[8], mul[8];
};

template <typename OutT>
__device__ __forceinline__ void encc_load_column(EnccColumn<OutT>& C, const uint32_t* desc, int c, int sw, int nc,
                                                 const uint32_t* first /* the staged words of the thread's first unit */) {
    const int BW = 2 * sw + nc + 1;
    constexpr uint32_t ONE = sizeof(OutT) == 2 ? 0x3F80u : 0x3F800000u;
    constexpr uint32_t FULL = sizeof(OutT) == 2 ? 0xFFFFu : 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t d = desc[c * 8 + k];
        const uint32_t kind = d >> 30, su = (d >> 27) & 7u, pl = d & 0x07FFFFFFu;
        uint32_t off = (uint32_t)(2 * sw + nc), sh = 0u, msk = 0u, mul = 0u;          // the tile's zero word
        if (kind == 1u) { off = pl >> 5; sh = pl & 31u; msk = 1u; mul = ONE; }
        else if (kind == 2u) { off = (uint32_t)sw + (pl >> 5); sh = pl & 31u; msk = 1u; mul = ONE; }
        else if (kind == 3u) { off = (uint32_t)(2 * sw) + pl; msk = FULL; mul = 1u; }
        if (sizeof(OutT) == 2 && (k & 1)) mul <<= 16;        // bf16: the odd element of a pair lands in the high half
        C.addr[k] = encc_addr(first + su * BW + off); C.sh[k] = sh; C.msk[k] = msk; C.mul[k] = mul;
    }
}

template <typename OutT>
__global__ void __launch_bounds__(ENCC_THREADS, 4)
encode_col_kernel(const __grid_constant__ EncParams E, int64_t n, int sw, const uint32_t* __restrict__ states,
                  const uint32_t* __restrict__ prev, const uint8_t* __restrict__ prev_count, int max_prev,
                  OutT* __restrict__ out, int c_pad, int row_elems, int S /* states per tile, a multiple of U */,
                  int U /* states per unit: U * row_elems is a multiple of 8 */) {
    ENCC_DYNAMIC_SMEM(csm);
    uint32_t* desc = csm;                                        // [U * row_elems]
    uint32_t* blk = desc + U * row_elems;                            // [S][BW]: state | OR of previous states | control values | 0
    const int nc = E.num_control_channels;
    const int BW = 2 * sw + nc + 1;
    const bool has_prev = prev && E.num_prev_states > 0;
    const int t = threadIdx.x;
    encc_build_desc(E, desc, row_elems, c_pad, has_prev, t, ENCC_THREADS, U);
    __syncthreads();
    const int unit_elems = U * row_elems;
    const int CPR = unit_elems >> 3;                             // 8-element chunks per unit
    const bool wide = CPR >= ENCC_THREADS;
    const int J = wide ? 1 : ENCC_THREADS / CPR;                 // units in flight per pass over the columns
    const int NT = wide ? ENCC_THREADS : J * CPR;                // threads in use
    const int NCG = wide ? (CPR + ENCC_THREADS - 1) / ENCC_THREADS : 1;
    const int j = wide ? 0 : t / CPR;
    const int c0 = wide ? t : t - j * CPR;
    EnccColumn<OutT> C;
    if (NCG == 1 && t < NT) encc_load_column<OutT>(C, desc, c0, sw, nc, blk + j * U * BW);
    const int step_s = ENCC_THREADS / sw, step_w = ENCC_THREADS - step_s * sw;     // (s, w) of word i + ENCC_THREADS
    for (int64_t n0 = (int64_t)blockIdx.x * S; n0 < n; n0 += (int64_t)gridDim.x * S) {
        const int nb = (int)((n - n0) < S ? (n - n0) : S);
        __syncthreads();                                         // the previous tile has been read
        {   // the tile's state words are one contiguous run in HBM; (state, word) advance without a division
            int s = t / sw, w = t - s * sw;
            const uint32_t* src = states + n0 * sw;
            for (int i = t; i < nb * sw; i += ENCC_THREADS) {
                uint32_t p = 0u;
                if (has_prev) {
                    const int np = prev_count ? prev_count[n0 + s] : max_prev;
                    for (int k = 0; k < np && k < max_prev; ++k) p |= prev[((n0 + s) * max_prev + k) * sw + w];
                }
                blk[s * BW + w] = src[i];
                blk[s * BW + sw + w] = p;
                s += step_s; w += step_w;
                if (w >= sw) { w -= sw; ++s; }
            }
        }
        for (int i = t; i < nb; i += ENCC_THREADS) blk[i * BW + 2 * sw + nc] = 0u;
        __syncthreads();
        for (int i = t; i < nb * nc; i += ENCC_THREADS) {
            const int s = i / nc, c = i - s * nc;
            const float v = enc_control_value(E, blk + s * BW, c);
            blk[s * BW + 2 * sw + c] =
                sizeof(OutT) == 2 ? (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v)) : __float_as_uint(v);
        }
        __syncthreads();
        const int nbu = nb / U;                                  // complete units of the tile
        if (t < NT) {
            for (int cg = 0; cg < NCG; ++cg) {
                const int c = cg * ENCC_THREADS + c0;
                if (c >= CPR) break;
                if (NCG > 1) encc_load_column<OutT>(C, desc, c, sw, nc, blk + j * U * BW);
                OutT* dst = out + n0 * (int64_t)row_elems + (int64_t)c * 8 + (int64_t)j * unit_elems;
                const int64_t dstep = (int64_t)J * unit_elems;
                const encc_addr_t astep = (encc_addr_t)(J * U * BW) * 4;
                int iters = 0;
                for (int u = j; u < nbu; u += J, ++iters) {
                    if (sizeof(OutT) == 2) {
                        uint32_t v[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k)            // two elements per word: the multipliers place them
                            v[k] = ((encc_lds(C.addr[2 * k]) >> C.sh[2 * k]) & C.msk[2 * k]) * C.mul[2 * k] +
                                   ((encc_lds(C.addr[2 * k + 1]) >> C.sh[2 * k + 1]) & C.msk[2 * k + 1]) * C.mul[2 * k + 1];
                        *reinterpret_cast<uint4*>(dst) = make_uint4(v[0], v[1], v[2], v[3]);
                    } else {
                        uint32_t v[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[k] = ((encc_lds(C.addr[k]) >> C.sh[k]) & C.msk[k]) * C.mul[k];
                        reinterpret_cast<uint4*>(dst)[0] = make_uint4(v[0], v[1], v[2], v[3]);
                        reinterpret_cast<uint4*>(dst)[1] = make_uint4(v[4], v[5], v[6], v[7]);
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) C.addr[k] += astep;
                    dst += dstep;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) C.addr[k] -= astep * (encc_addr_t)iters;      // back to the first unit (next tile)
            }
        }
        // states beyond the last complete unit (the end of a batch that is not a multiple of U): element by element
        for (int e = t; e < (nb - nbu * U) * row_elems; e += ENCC_THREADS) {
            constexpr uint32_t ONE = sizeof(OutT) == 2 ? 0x3F80u : 0x3F800000u;
            const int sr = e / row_elems, le = e - sr * row_elems, s = nbu * U + sr;
            const uint32_t d = desc[le];
            const uint32_t kind = d >> 30, pl = d & 0x07FFFFFFu;
            const uint32_t* b = blk + s * BW;
            uint32_t x = 0u;
            if (kind == 1u) x = ((b[pl >> 5] >> (pl & 31u)) & 1u) ? ONE : 0u;
            else if (kind == 2u) x = ((b[sw + (pl >> 5)] >> (pl & 31u)) & 1u) ? ONE : 0u;
            else if (kind == 3u) x = b[2 * sw + pl];
            const int64_t o = (n0 + s) * (int64_t)row_elems + le;
            if (sizeof(OutT) == 2) reinterpret_cast<uint16_t*>(out)[o] = (uint16_t)x;
            else reinterpret_cast<uint32_t*>(out)[o] = x;
        }
    }
}

#endif  // GZ_ENCODER_COL_CUH
