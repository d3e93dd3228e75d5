/* gz_propnet_format.h — flat int32 "blob" describing one compiled propnet.
 *
 * Produced by galvanise_zero_b200/tables.py (pack_blob), consumed by the CUDA library
 * (csrc/gz_capi.cu) and by the CPU oracle (oracle/propnet_ref.c). Plain data: no pointers.
 *
 * It replaces what the reference obtains from ggplib when it does
 *   sm = lookup.by_name(game).get_sm()                       (src/ggpzero/nn/manager.py:62-70,
 *                                                             src/test/cpp/test_interface.py:40-44)
 * i.e. the per-game propositional network that StateMachineInterface evaluates.
 *
 * Literal encoding: lit = slot*2 + inverted.  Slot 0 is constant TRUE, slots 1..n_bases are the
 * base propositions, then the inputs role-major, then gate outputs. Every gate is an AND over its
 * literals (OR is expressed through inverted literals / inverted references).
 *
 * Packed state convention: bit b of a state lives in 32-bit word b>>5 at bit b&31 (LSB first);
 * a state row is `state_words` words (padded to a multiple of 4 words = 16 bytes).
 * Packed legal row: for each role r, ceil(A_r/32) words, concatenated (offset legal_word_off[r]);
 * row padded to a multiple of 4 words (`legal_words`).
 */
#ifndef GZ_PROPNET_FORMAT_H
#define GZ_PROPNET_FORMAT_H

#include <stdint.h>

#define GZ_PN_MAGIC 0x475A504E /* 'GZPN' */
#define GZ_PN_VERSION 3
#define GZ_PN_HEADER_WORDS 48
#define GZ_PN_MAX_ROLES 4

/* header word indices */
enum {
    GZ_H_MAGIC = 0,
    GZ_H_VERSION = 1,
    GZ_H_N_BASES = 2,
    GZ_H_N_ROLES = 3,
    GZ_H_N_SLOTS = 4,
    GZ_H_N_GATES = 5,
    GZ_H_TERMINAL_LIT = 6,
    GZ_H_STATE_WORDS = 7,   /* padded words per packed state row */
    GZ_H_LEGAL_WORDS = 8,   /* padded words per packed legal row */
    GZ_H_N_INPUTS = 9,      /* sum of role sizes */
    GZ_H_TOTAL_WORDS = 10,
    /* section offsets (in int32 words from blob start) */
    GZ_H_OFF_ROLE_SIZES = 11,   /* [n_roles] */
    GZ_H_OFF_GATE_OUT = 12,     /* [n_gates] slot written by gate i */
    GZ_H_OFF_LIT_OFF = 13,      /* [n_gates+1] */
    GZ_H_OFF_LIT_DATA = 14,     /* [lit_off[n_gates]] */
    GZ_H_OFF_NEXT_LIT = 15,     /* [n_bases] */
    GZ_H_OFF_LEGAL_OFF = 16,    /* [n_roles+1] offsets into legal_lit */
    GZ_H_OFF_LEGAL_LIT = 17,    /* [n_inputs] */
    GZ_H_OFF_GOAL_OFF = 18,     /* [n_roles+1] */
    GZ_H_OFF_GOAL_VAL = 19,     /* [n_goal] */
    GZ_H_OFF_GOAL_LIT = 20,     /* [n_goal] */
    GZ_H_OFF_INIT = 21,         /* [state_words] packed initial state */
    GZ_H_OFF_PROG_V = 22,  GZ_H_LEN_PROG_V = 23,   /* gate indices, level order */
    GZ_H_OFF_PROG_T = 24,  GZ_H_LEN_PROG_T = 25,
    GZ_H_OFF_PROG_TV = 26, GZ_H_LEN_PROG_TV = 27,
    GZ_H_OFF_PROG_LT = 28, GZ_H_LEN_PROG_LT = 29,
    GZ_H_OFF_PROG_GL = 30, GZ_H_LEN_PROG_GL = 31,
    GZ_H_OFF_LEGAL_WORD_OFF = 32, /* [n_roles+1] word offset of each role inside a legal row */
    GZ_H_OFF_PROG_TL = 33, GZ_H_LEN_PROG_TL = 34   /* T minus LT */
};

/* Counter-based RNG shared bit-for-bit by the oracle and the CUDA playout kernels:
 * pick for (seed, slot, step, role); index = mulhi32(r, count). */
#if defined(__CUDACC__)
#define GZ_HD __host__ __device__ __forceinline__
#else
#define GZ_HD static inline
#endif

GZ_HD uint32_t gz_rand32(uint64_t seed, uint64_t slot, uint64_t step, uint32_t role) {
    uint64_t z = seed ^ (slot * 0x9E3779B97F4A7C15ULL) ^ ((step + 1) * 0xBF58476D1CE4E5B9ULL)
                 ^ ((uint64_t)(role + 1) * 0x94D049BB133111EBULL);
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}

GZ_HD uint32_t gz_pick(uint32_t r, uint32_t count) {
    return (uint32_t)(((uint64_t)r * (uint64_t)count) >> 32);
}

#endif
