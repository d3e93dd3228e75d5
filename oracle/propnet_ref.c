/* ORACLE — test infrastructure. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path never does.
 *
 * Plain-C, one-state-at-a-time restatement of the propnet state machine the reference calls through
 * ggplib's GGPLib::StateMachineInterface (not vendored in the reference, no version pinned:
 * src/cpp/Makefile:1-7). The call protocol follows the reference's own call sites:
 *
 *   updateBases(state)                  src/cpp/puct/node.cpp:157, src/cpp/puct/evaluator.cpp:218
 *   getLegalState(ri)->getCount/getLegal src/cpp/puct/node.cpp:115-120,170-188
 *   isTerminal / getGoalValue(ri)       src/cpp/puct/node.cpp:162,214
 *   nextState(joint_move, out)          src/cpp/puct/evaluator.cpp:219
 *   random playout loop                 src/cpp/pyobjects/gdltransformer_impl.cpp:50-124
 *
 * Algorithm restated: forward propagation of a propositional network (AND/OR/NOT components,
 * transitions read after propagation) — the published propnet state-machine scheme of
 * ggp-base / ggplib. The gate tables come from include/gz_propnet_format.h blobs.
 *
 * PARITY STATUS: "parity unpinned" against ggplib itself (absent). Pinned instead against the
 * from-rules GDL interpreter oracle/gdl_interp.py through tests/golden/ fixtures.
 */
#define _POSIX_C_SOURCE 200809L
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>

#include "../include/gz_propnet_format.h"

typedef struct {
    int n;          /* gates */
    int32_t* out;   /* [n] */
    int32_t* off;   /* [n+1] */
    int32_t* lits;  /* flattened */
} pnref_prog;

typedef struct pnref_sm {
    int32_t* blob;
    int64_t n_words;
    const int32_t* h;
    int n_bases, n_roles, n_slots, n_gates, state_words, legal_words, n_inputs;
    const int32_t *role_sizes, *next_lit, *legal_off, *legal_lit, *goal_off, *goal_val, *goal_lit,
        *init, *legal_word_off;
    pnref_prog V, T, TV, LT, GL, TL;
    uint8_t* val;        /* [n_slots] */
    int32_t* legal_list; /* [n_inputs] ascending legal indices per role */
    int32_t legal_count[GZ_PN_MAX_ROLES];
    int goals_valid; /* the goal cone (GL) is propagated lazily, on the first getGoalValue */
} pnref_sm;

static void build_prog(pnref_sm* sm, pnref_prog* p, int off_idx, int len_idx) {
    const int32_t* h = sm->h;
    const int32_t* idx = sm->blob + h[off_idx];
    const int32_t* gout = sm->blob + h[GZ_H_OFF_GATE_OUT];
    const int32_t* loff = sm->blob + h[GZ_H_OFF_LIT_OFF];
    const int32_t* ldat = sm->blob + h[GZ_H_OFF_LIT_DATA];
    int n = h[len_idx];
    int64_t tot = 0;
    for (int i = 0; i < n; i++) tot += loff[idx[i] + 1] - loff[idx[i]];
    p->n = n;
    p->out = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
    p->off = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
    p->lits = (int32_t*)malloc(sizeof(int32_t) * (tot + 1));
    int32_t pos = 0;
    for (int i = 0; i < n; i++) {
        int g = idx[i];
        p->out[i] = gout[g];
        p->off[i] = pos;
        for (int k = loff[g]; k < loff[g + 1]; k++) p->lits[pos++] = ldat[k];
    }
    p->off[n] = pos;
}

static void free_prog(pnref_prog* p) {
    free(p->out); free(p->off); free(p->lits);
}

static inline void run_prog(const pnref_prog* p, uint8_t* val) {
    const int32_t* lits = p->lits;
    for (int i = 0; i < p->n; i++) {
        uint8_t acc = 1;
        for (int k = p->off[i]; k < p->off[i + 1]; k++) {
            int32_t l = lits[k];
            acc &= (uint8_t)(val[l >> 1] ^ (l & 1));
        }
        val[p->out[i]] = acc;
    }
}

static inline int lit_val(const uint8_t* val, int32_t l) {
    return (val[l >> 1] ^ (l & 1)) & 1;
}

pnref_sm* pnref_create(const int32_t* blob, int64_t n_words) {
    if (n_words < GZ_PN_HEADER_WORDS || blob[GZ_H_MAGIC] != GZ_PN_MAGIC ||
        blob[GZ_H_VERSION] != GZ_PN_VERSION || blob[GZ_H_TOTAL_WORDS] != n_words ||
        blob[GZ_H_N_ROLES] > GZ_PN_MAX_ROLES)
        return NULL;
    pnref_sm* sm = (pnref_sm*)calloc(1, sizeof(pnref_sm));
    sm->blob = (int32_t*)malloc(sizeof(int32_t) * n_words);
    memcpy(sm->blob, blob, sizeof(int32_t) * n_words);
    sm->n_words = n_words;
    const int32_t* h = sm->h = sm->blob;
    sm->n_bases = h[GZ_H_N_BASES];
    sm->n_roles = h[GZ_H_N_ROLES];
    sm->n_slots = h[GZ_H_N_SLOTS];
    sm->n_gates = h[GZ_H_N_GATES];
    sm->state_words = h[GZ_H_STATE_WORDS];
    sm->legal_words = h[GZ_H_LEGAL_WORDS];
    sm->n_inputs = h[GZ_H_N_INPUTS];
    sm->role_sizes = sm->blob + h[GZ_H_OFF_ROLE_SIZES];
    sm->next_lit = sm->blob + h[GZ_H_OFF_NEXT_LIT];
    sm->legal_off = sm->blob + h[GZ_H_OFF_LEGAL_OFF];
    sm->legal_lit = sm->blob + h[GZ_H_OFF_LEGAL_LIT];
    sm->goal_off = sm->blob + h[GZ_H_OFF_GOAL_OFF];
    sm->goal_val = sm->blob + h[GZ_H_OFF_GOAL_VAL];
    sm->goal_lit = sm->blob + h[GZ_H_OFF_GOAL_LIT];
    sm->init = sm->blob + h[GZ_H_OFF_INIT];
    sm->legal_word_off = sm->blob + h[GZ_H_OFF_LEGAL_WORD_OFF];
    build_prog(sm, &sm->V, GZ_H_OFF_PROG_V, GZ_H_LEN_PROG_V);
    build_prog(sm, &sm->T, GZ_H_OFF_PROG_T, GZ_H_LEN_PROG_T);
    build_prog(sm, &sm->TV, GZ_H_OFF_PROG_TV, GZ_H_LEN_PROG_TV);
    build_prog(sm, &sm->LT, GZ_H_OFF_PROG_LT, GZ_H_LEN_PROG_LT);
    build_prog(sm, &sm->GL, GZ_H_OFF_PROG_GL, GZ_H_LEN_PROG_GL);
    build_prog(sm, &sm->TL, GZ_H_OFF_PROG_TL, GZ_H_LEN_PROG_TL);
    sm->val = (uint8_t*)calloc(sm->n_slots + 8, 1);
    sm->val[0] = 1;
    sm->legal_list = (int32_t*)malloc(sizeof(int32_t) * (sm->n_inputs + 1));
    return sm;
}

void pnref_destroy(pnref_sm* sm) {
    if (!sm) return;
    free_prog(&sm->V); free_prog(&sm->T); free_prog(&sm->TV); free_prog(&sm->LT); free_prog(&sm->GL); free_prog(&sm->TL);
    free(sm->val); free(sm->legal_list); free(sm->blob); free(sm);
}

pnref_sm* pnref_dupe(const pnref_sm* sm) {
    pnref_sm* d = pnref_create(sm->blob, sm->n_words);
    memcpy(d->val, sm->val, sm->n_slots);
    memcpy(d->legal_list, sm->legal_list, sizeof(int32_t) * sm->n_inputs);
    memcpy(d->legal_count, sm->legal_count, sizeof(sm->legal_count));
    d->goals_valid = sm->goals_valid;
    return d;
}

int pnref_n_bases(const pnref_sm* sm) { return sm->n_bases; }
int pnref_n_roles(const pnref_sm* sm) { return sm->n_roles; }
int pnref_role_size(const pnref_sm* sm, int r) { return sm->role_sizes[r]; }
int pnref_state_words(const pnref_sm* sm) { return sm->state_words; }
int pnref_legal_words(const pnref_sm* sm) { return sm->legal_words; }
void pnref_initial_state(const pnref_sm* sm, uint32_t* out) {
    memcpy(out, sm->init, sizeof(uint32_t) * sm->state_words);
}

static void load_bases(pnref_sm* sm, const uint32_t* state) {
    for (int b = 0; b < sm->n_bases; b++) sm->val[1 + b] = (state[b >> 5] >> (b & 31)) & 1;
}

/* updateBases: load the base bits, propagate the views (legal / terminal / goal cone). */
void pnref_update_bases(pnref_sm* sm, const uint32_t* state) {
    load_bases(sm, state);
    run_prog(&sm->LT, sm->val);
    sm->goals_valid = 0;
    for (int r = 0; r < sm->n_roles; r++) {
        int32_t* list = sm->legal_list + sm->legal_off[r];
        int c = 0;
        for (int a = 0; a < sm->role_sizes[r]; a++)
            if (lit_val(sm->val, sm->legal_lit[sm->legal_off[r] + a])) list[c++] = a;
        sm->legal_count[r] = c;
    }
}

int pnref_legal_count(const pnref_sm* sm, int role) { return sm->legal_count[role]; }
int pnref_legal(const pnref_sm* sm, int role, int i) {
    return sm->legal_list[sm->legal_off[role] + i];
}
int pnref_is_terminal(const pnref_sm* sm) { return lit_val(sm->val, sm->h[GZ_H_TERMINAL_LIT]); }

/* goal value of `role`: the largest value whose goal proposition holds; -1 when none does. */
int pnref_goal(pnref_sm* sm, int role) {
    int res = -1;
    if (!sm->goals_valid) {
        run_prog(&sm->GL, sm->val);
        sm->goals_valid = 1;
    }
    for (int k = sm->goal_off[role]; k < sm->goal_off[role + 1]; k++)
        if (lit_val(sm->val, sm->goal_lit[k])) res = sm->goal_val[k];
    return res;
}

/* nextState: applies one action index per role to the bases loaded by the last updateBases. */
void pnref_next_state(pnref_sm* sm, const int32_t* joint_move, uint32_t* out) {
    uint8_t* in = sm->val + 1 + sm->n_bases;
    memset(in, 0, sm->n_inputs);
    int off = 0;
    for (int r = 0; r < sm->n_roles; r++) {
        in[off + joint_move[r]] = 1;
        off += sm->role_sizes[r];
    }
    /* only the legal/terminal cone is known to be current: run the rest of the next-state cone.
       TL may overwrite goal-only gates it shares with GL with the same values. */
    run_prog(&sm->TL, sm->val);
    memset(out, 0, sizeof(uint32_t) * sm->state_words);
    for (int b = 0; b < sm->n_bases; b++)
        if (lit_val(sm->val, sm->next_lit[b])) out[b >> 5] |= 1u << (b & 31);
}

/* ---------------------------------------------------------------------------- batch helpers */
static void write_views(pnref_sm* sm, uint32_t* legal_row, uint8_t* term, uint8_t* goals) {
    memset(legal_row, 0, sizeof(uint32_t) * sm->legal_words);
    for (int r = 0; r < sm->n_roles; r++) {
        uint32_t* w = legal_row + sm->legal_word_off[r];
        for (int i = 0; i < sm->legal_count[r]; i++) {
            int a = sm->legal_list[sm->legal_off[r] + i];
            w[a >> 5] |= 1u << (a & 31);
        }
        int g = pnref_goal(sm, r);
        goals[r] = (uint8_t)(g < 0 ? 255 : g);
    }
    *term = (uint8_t)pnref_is_terminal(sm);
}

/* states [n][state_words] -> legal [n][legal_words], terminal [n], goals [n][n_roles] */
void pnref_eval_batch(pnref_sm* sm, int64_t n, const uint32_t* states, uint32_t* legal,
                      uint8_t* terminal, uint8_t* goals) {
    for (int64_t i = 0; i < n; i++) {
        pnref_update_bases(sm, states + i * sm->state_words);
        write_views(sm, legal + i * sm->legal_words, terminal + i, goals + i * sm->n_roles);
    }
}

/* one transition per state: next = nextState(state, moves); outputs describe `next`. */
void pnref_transition_batch(pnref_sm* sm, int64_t n, const uint32_t* states, const int16_t* moves,
                            uint32_t* next, uint32_t* legal, uint8_t* terminal, uint8_t* goals) {
    int32_t jm[GZ_PN_MAX_ROLES];
    for (int64_t i = 0; i < n; i++) {
        pnref_update_bases(sm, states + i * sm->state_words);
        for (int r = 0; r < sm->n_roles; r++) jm[r] = moves[i * sm->n_roles + r];
        uint32_t* o = next + i * sm->state_words;
        pnref_next_state(sm, jm, o);
        pnref_update_bases(sm, o);
        write_views(sm, legal + i * sm->legal_words, terminal + i, goals + i * sm->n_roles);
    }
}

/* k-th (0-based) set bit of a packed mask */
static int kth_set(const uint32_t* w, int nbits, uint32_t k) {
    for (int i = 0; i < (nbits + 31) / 32; i++) {
        uint32_t x = w[i];
        uint32_t c = (uint32_t)__builtin_popcount(x);
        if (k < c) {
            while (k--) x &= x - 1;
            return i * 32 + __builtin_ctz(x);
        }
        k -= c;
    }
    return -1;
}

/* One random-playout step for slots [slot0, slot0+n): terminal slots restart from the initial state,
 * others pick a uniformly random legal action per role with gz_rand32(seed, slot, step, role) and
 * advance. In/out arrays describe the current state of each slot. moves_out (optional) records the
 * chosen actions (-1 for a restart). Returns the number of real transitions. */
int64_t pnref_playout_step(pnref_sm* sm, int64_t slot0, int64_t n, uint64_t seed, uint64_t step,
                           uint32_t* states, uint32_t* legal, uint8_t* terminal, uint8_t* goals,
                           int16_t* moves_out) {
    int64_t count = 0;
    int32_t jm[GZ_PN_MAX_ROLES];
    for (int64_t i = 0; i < n; i++) {
        uint32_t* s = states + i * sm->state_words;
        uint32_t* lg = legal + i * sm->legal_words;
        if (terminal[i]) {
            memcpy(s, sm->init, sizeof(uint32_t) * sm->state_words);
            for (int r = 0; r < sm->n_roles; r++)
                if (moves_out) moves_out[i * sm->n_roles + r] = -1;
            pnref_update_bases(sm, s);
        } else {
            for (int r = 0; r < sm->n_roles; r++) {
                const uint32_t* w = lg + sm->legal_word_off[r];
                uint32_t c = 0;
                for (int k = 0; k < (sm->role_sizes[r] + 31) / 32; k++)
                    c += (uint32_t)__builtin_popcount(w[k]);
                uint32_t pick = gz_pick(gz_rand32(seed, (uint64_t)(slot0 + i), step, (uint32_t)r), c);
                jm[r] = kth_set(w, sm->role_sizes[r], pick);
                if (moves_out) moves_out[i * sm->n_roles + r] = (int16_t)jm[r];
            }
            pnref_update_bases(sm, s);
            pnref_next_state(sm, jm, s);
            pnref_update_bases(sm, s);
            count++;
        }
        write_views(sm, lg, terminal + i, goals + i * sm->n_roles);
    }
    return count;
}

/* CPU baseline: `n_slots` concurrent random playouts for `n_steps` steps on `n_threads` threads,
 * following the reference's own playout loop (gdltransformer_impl.cpp:50-124: per ply read the legal
 * sets, pick uniformly per role, nextState, updateBases). Returns seconds; *transitions = total. */
typedef struct {
    const int32_t* blob;
    int64_t n_words, lo, hi, n_steps, total;
    uint64_t seed, cks;
} pnref_job;

static void* playout_worker(void* arg) {
    pnref_job* j = (pnref_job*)arg;
    pnref_sm* sm = pnref_create(j->blob, j->n_words);
    uint32_t* st = (uint32_t*)calloc((size_t)sm->state_words + 4, 4);
    int32_t jm[GZ_PN_MAX_ROLES];
    int64_t total = 0;
    uint64_t cks = 0;
    /* slot-major: the state machine keeps its propagated views between plies, so one ply costs
     * nextState + updateBases exactly as in the reference loop. Same (slot, step) RNG stream as
     * pnref_playout_step, hence the same trajectories. */
    for (int64_t slot = j->lo; slot < j->hi; slot++) {
        int term = 1;
        for (int64_t s = 0; s < j->n_steps; s++) {
            if (term) {
                memcpy(st, sm->init, sizeof(uint32_t) * sm->state_words);
            } else {
                for (int r = 0; r < sm->n_roles; r++) {
                    uint32_t pick = gz_pick(gz_rand32(j->seed, (uint64_t)slot, (uint64_t)s, (uint32_t)r),
                                            (uint32_t)sm->legal_count[r]);
                    jm[r] = sm->legal_list[sm->legal_off[r] + pick];
                }
                pnref_next_state(sm, jm, st);
                total++;
            }
            pnref_update_bases(sm, st);
            term = pnref_is_terminal(sm);
        }
        for (int i = 0; i < sm->state_words; i++)
            cks ^= (uint64_t)st[i] * 0x9E3779B97F4A7C15ULL + (uint64_t)(slot * sm->state_words + i);
    }
    free(st);
    pnref_destroy(sm);
    j->total = total;
    j->cks = cks;
    return NULL;
}

double pnref_playout_bench(const int32_t* blob, int64_t n_words, int64_t n_slots, int64_t n_steps,
                           uint64_t seed, int n_threads, int64_t* transitions, uint64_t* checksum) {
    struct timespec t0, t1;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    pnref_job* jobs = (pnref_job*)calloc((size_t)n_threads, sizeof(pnref_job));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < n_threads; t++) {
        jobs[t].blob = blob; jobs[t].n_words = n_words; jobs[t].n_steps = n_steps; jobs[t].seed = seed;
        jobs[t].lo = n_slots * t / n_threads; jobs[t].hi = n_slots * (t + 1) / n_threads;
        pthread_create(&th[t], NULL, playout_worker, &jobs[t]);
    }
    int64_t total = 0;
    uint64_t cks = 0;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], NULL);
        total += jobs[t].total;
        cks ^= jobs[t].cks;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(jobs); free(th);
    if (transitions) *transitions = total;
    if (checksum) *checksum = cks;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
