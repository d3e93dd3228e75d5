// Memory layout of a PuctEngine: one arena, every array 256-byte aligned. Used by gz_puct.cu
// (cudaMalloc'd arena) and by the CPU validation build in tests/hostsim (malloc'd arena).
#ifndef GZ_PUCT_LAYOUT_H
#define GZ_PUCT_LAYOUT_H

#include <stddef.h>
#include <stdint.h>

#include "gz_puct_core.h"

// Assigns all array pointers of E inside [base, base+return). With base == nullptr only the size
// is computed. Dimensions (G, EB, R, SW, LW, NCAP, CCAP, MAXD, MAXPEND, srec, READYCAP, uniq_cap, tt_cap,
// n_rewards, policy_stride) must be set beforehand.
static inline size_t gzp_layout(PuctEngine* E, uint8_t* base) {
    size_t pos = 0;
    auto take = [&](size_t bytes) -> uint8_t* {
        uint8_t* p = base ? base + pos : nullptr;
        pos += (bytes + 255) & ~(size_t)255;
        return p;
    };
    const size_t NN = (size_t)E->G * E->NCAP;
    const size_t CC = (size_t)E->G * E->CCAP;
    const size_t G = (size_t)E->G;
    const size_t EB = (size_t)(E->EB < 1 ? 1 : E->EB);
    E->nodes = (GzpNode*)take(NN * sizeof(GzpNode));
    E->n_state = (uint32_t*)take(NN * 4 * E->SW);
    E->n_other = (int32_t*)take(NN * 4 * GZP_MAX_ROLES);
    E->children = (GzpChild*)take(CC * sizeof(GzpChild));
    E->c_owner = (int32_t*)take(CC * 4);
    E->n_newidx = (int32_t*)take(NN * 4);
    E->n_newc0 = (int32_t*)take(NN * 4);
    E->ctl = (GameCtl*)take(G * sizeof(GameCtl));
    E->path_node = (int32_t*)take(G * E->MAXD * 4);
    E->path_choice = (int32_t*)take(G * E->MAXD * 4);
    E->root_prev_state = (uint32_t*)take(G * E->SW * 4);
    E->x_in_state = (uint32_t*)take(G * E->SW * 4);
    E->x_state = (uint32_t*)take(G * E->SW * 4);
    E->x_legal = (uint32_t*)take(G * E->LW * 4);
    E->x_moves = (int16_t*)take(G * E->R * 2);
    E->x_term = (uint8_t*)take(G);
    E->x_goals = (uint8_t*)take(G * E->R);
    E->x_prevcnt = (uint8_t*)take(G);
    E->x_need_eval = (uint8_t*)take(G * EB);
    if (EB > 1) {
        // intra-tree workers: EB evaluation rows per game and the coroutines' queues / paths
        E->e_state = (uint32_t*)take(G * EB * E->SW * 4);
        E->e_prev = (uint32_t*)take(G * EB * E->SW * 4);
        E->e_prevcnt = (uint8_t*)take(G * EB);
        E->sched = (GzpSched*)take(G * sizeof(GzpSched));
        E->co_path_node = (int32_t*)take(G * (EB + 1) * E->MAXD * 4);
        E->co_path_choice = (int32_t*)take(G * (EB + 1) * E->MAXD * 4);
    } else {
        E->e_state = E->x_state;
        E->e_prev = E->x_in_state;
        E->e_prevcnt = E->x_prevcnt;
        E->sched = nullptr;
        E->co_path_node = E->path_node;
        E->co_path_choice = E->path_choice;
    }
    E->x_cache = (uint32_t*)take(G * gzp_cache_words(E->max_actions) * 4);
    E->pend = (int32_t*)take(G * E->MAXPEND * (size_t)E->srec * 4);
    E->ready = (int32_t*)take((size_t)E->READYCAP * E->srec * 4);
    E->ready_count = (unsigned long long*)take(16);
    E->uniq_keys = (unsigned long long*)take((size_t)E->uniq_cap * 8);
    E->uniq_counts = (int32_t*)take((size_t)E->uniq_cap * 4);
    E->uniq_mask = (const uint32_t*)take((size_t)E->SW * 4);
    E->init_state = (const uint32_t*)take((size_t)E->SW * 4);
    E->start_state = (uint32_t*)take(G * E->SW * 4);
    E->tt_keys = (unsigned long long*)take(G * (size_t)E->tt_cap * 8);      // tt_cap == 0: no transposition table
    E->tt_node = (int32_t*)take(G * (size_t)E->tt_cap * 4);
    return pos;
}

static inline int gzp_sample_record_words(int SW, int max_actions) {
    return SR_HEADER + 2 * SW + 2 * max_actions;
}

#endif
