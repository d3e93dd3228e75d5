// gz_puct_core.h — PUCT search + self-play control for ONE game tree, written as host/device code.
//
// The same source is compiled into the CUDA kernels of gz_puct.cu (one game per warp; lane 0 runs
// the sequential control flow) and, for CPU-side validation only, into tests/hostsim (g++), where
// it is driven against the reference's own C++ compiled in oracle/_ref.
//
// It re-states, batch-synchronously, what the reference does with one coroutine per game:
//   PuctNode::create / createNode        src/cpp/puct/node.cpp:42-221, evaluator.cpp:165-214
//   PuctNodeRequest::reply               src/cpp/puct/node.cpp:463-511
//   selectChild / priorScore / setPuctConstant / setDirichletNoise
//                                        src/cpp/puct/evaluator.cpp:341-517,1195-1307
//   backup                               src/cpp/puct/evaluator.cpp:519-656
//   treePlayout / playoutMain / onNextMove / converged
//                                        src/cpp/puct/evaluator.cpp:658-886,1032-1098,1342-1362
//   choose / chooseTopVisits / chooseTemperature / getProbabilities / getTemperature
//                                        src/cpp/puct/evaluator.cpp:1100-1192,1309-1340,1473-1510
//   fastApplyMove / resetRootNode / establishRoot     evaluator.cpp:888-1030
//   SelfPlay::playOnce / collectSamples / runToEnd / resign / addSamples   src/cpp/selfplay.cpp:45-343
//   SelfPlayManager::createSample, UniqueStates       selfplaymanager.cpp:72-119, uniquestates.h:12-79
// Arithmetic types (float vs double, rounding points) follow the reference expression by
// expression so that, with identical network outputs, visit counts are identical.
//
// Scope notes: games in which two roles have a real choice in the same state are rejected. Self-play runs one
// evaluation per tree at a time (PuctConfig.batch_size == 1, templates.py:76); a Player's search with
// batch_size > 1 runs the reference's batch_size - 1 playoutWorker coroutines beside playoutMain
// (evaluator.cpp:722-742,1050-1069) as resumable state machines under a restatement of NetworkScheduler's
// run queue (scheduler.cpp:131-203), see "Intra-tree workers" below.
// With PuctConfig.lookup_transpositions the trees are DAGs (a node reached by several move orders is shared,
// evaluator.cpp:216-239): the copy of a child node's statistics kept in its parent's child record (GzpChild)
// would go stale in the parents a backup does not pass through, so select then reads the child nodes themselves.
#ifndef GZ_PUCT_CORE_H
#define GZ_PUCT_CORE_H

#include <math.h>
#include <stdint.h>

#include "gz_libm.h"

#ifdef __CUDACC__
#define GZP_HD __host__ __device__ __forceinline__
#define GZP_HDN __host__ __device__
#define GZP_MEMBER __host__ __device__ __forceinline__
#define GZP_RARE __host__ __device__ __noinline__      // off the hot path: one out-of-line copy keeps the kernel's code small
#else
#define GZP_RARE static __attribute__((noinline))
#define GZP_MEMBER inline
#define GZP_HD static inline
#define GZP_HDN static inline
#endif

#define GZP_MAX_ROLES 4
#ifdef __CUDACC__
#define GZP_ALIGN(n) __align__(n)
#else
#define GZP_ALIGN(n) __attribute__((aligned(n)))
#endif

// ---- phases of the per-game controller
enum {
    PH_NEW_GAME = 0,
    PH_COLLECT_NEXT = 1,   // top of SelfPlay::collectSamples loop
    PH_SEARCH = 2,         // inside playoutMain
    PH_MOVE_DONE = 3,      // search finished: choose, sample, apply
    PH_AFTER_MOVE = 4,     // new root in place: post-move checks
    PH_RUNTOEND_NEXT = 5,  // top of SelfPlay::runToEnd loop
    PH_GAME_END = 6,
    PH_IDLE = 7,           // (player mode) waiting for a command
    PH_HALT = 8            // error
};

// ---- requests to the batch pipeline
enum { REQ_NONE = 0, REQ_EXPAND = 1, REQ_ROOT = 2, REQ_REROOT = 3 };
enum { PUR_PLAYOUT = 0, PUR_REROOT = 1, PUR_ROOT = 2 };

// ---- node flags
enum { NF_FINALISED = 1, NF_NOISE_SET = 2, NF_FORCE_TERMINAL = 4 };

enum { GZP_ERR_NONE = 0, GZP_ERR_NODES = 1, GZP_ERR_CHILDREN = 2, GZP_ERR_PATH = 3, GZP_ERR_SIMULTANEOUS = 4,
       GZP_ERR_SAMPLES = 5, GZP_ERR_ROOT_TERMINAL = 6 };

struct PuctConf {
    float puct_constant, puct_constant_root;
    float dirichlet_noise_pct, noise_policy_squash_pct, noise_policy_squash_prob;
    int choose;  // 0 top visits, 1 temperature
    float random_scale, temperature;
    int depth_temperature_start;
    float depth_temperature_increment;
    int depth_temperature_stop;
    float depth_temperature_max;
    float fpu_prior_discount, fpu_prior_discount_root;
    float top_visits_best_guess_converge_ratio;
    float think_time;
    int converged_visits;
    int batch_size;
    int use_legals_count_draw;
    int backup_finalised;
    int lookup_transpositions;
    float evaluation_multiplier_to_convergence;
};

struct SpConf {
    float oscillate_sampling_pct, temperature_for_policy;
    int evals_per_move;
    float resign0_score_probability, resign0_pct, resign1_score_probability, resign1_pct;
    float run_to_end_pct;
    int run_to_end_evals;
    float run_to_end_early_score;
    int run_to_end_minimum_game_depth;
    int abort_max_length;
    PuctConf puct;
    PuctConf run_to_end;
};

// ---- intra-tree workers (PuctConfig.batch_size > 1): one record per coroutine of a search, and the per-game
// run queue that replaces NetworkScheduler::mainLoop for them
#define GZP_MAX_EB 64
enum {
    CO_DEAD = 0,
    CO_MAIN_START,        // onNextMove: spawn the workers (evaluator.cpp:1050-1065)
    CO_MAIN_TOP,          // top of playoutMain's loop (evaluator.cpp:782)
    CO_MAIN_COLLECT,      // "collect workers" (evaluator.cpp:1076-1079)
    CO_WORKER_TOP,        // top of playoutWorker's loop (evaluator.cpp:723)
    CO_WORKER_RESUMED,    // ... after its yield
    CO_DESCEND,           // inside treePlayout's loop at node `cur`
    CO_SELECT_RESUMED,    // after the yield inside selectChild (bad fallback with unselectable siblings)
    CO_AFTER_PROPNET,     // expandChild: waiting for the propnet pass / node creation
    CO_AFTER_EVAL,        // createNode: blocked in NetworkScheduler::evaluate
    // balanceFirstMoves (evaluator.cpp:282-339): the first moves of the root are searched until they have as many
    // traversals as the most visited one
    CO_BAL_START, CO_BAL_NEXT, CO_BAL_LOOP, CO_BAL_WAIT,
    CO_BAL_WORKER         // one playout below a given root child, then done
};
enum { CO_KIND_LOOP = 0, CO_KIND_ONCE = 1 };       // after a playout: back to the top of the loop / finished
struct GzpCo {
    int32_t st;
    int32_t cur, path_len;
    int32_t req_parent, req_child;   // the child it is expanding (that child is `unselectable` meanwhile)
    int32_t node;                    // the node created for it
    int32_t sel_to;                  // CO_SELECT_RESUMED: the node the chosen child leads to
    int32_t evaluated;               // the node went to the network (counts as an evaluation when the coroutine resumes)
    int32_t kind;                    // CO_KIND_*
    int32_t pad[3];
};
struct GzpSched {
    int32_t n_run, n_req, n_yld;     // runnables (FIFO), requestors, yielders
    int32_t current;                 // coroutine that owns the controller (it may be waiting for the propnet pass), -1
    int32_t worker_count, do_playouts, flush_wait, active;
    int32_t run[GZP_MAX_EB + 2], req[GZP_MAX_EB + 2], yld[GZP_MAX_EB + 2];
    GzpCo co[GZP_MAX_EB + 1];        // 0: the search itself (onNextMove / playoutMain / balanceFirstMoves), 1..: workers
    // balanceFirstMoves
    int32_t balance, bal_max, bal_i, bal_wanted;
    int32_t bal_order[32];           // the root's children by traversals (first bal_max of them)
};

// ---- per-game controller record
struct GameCtl {
    int32_t phase, arena, root, n_nodes, n_children, game_depth;
    uint64_t rng_eval[2], rng_sp[2];
    int32_t conf_sel, max_evaluations, num_evaluations, num_tree_playouts, playouts_finals;
    int32_t cur, path_len, in_playout;
    int32_t req, req_parent, req_child, req_purpose;
    int32_t eval_node, new_node, choice, resume;
    // player-mode command mailbox (mode == 1)
    int32_t cmd, cmd_arg, cmd_evals, cmd_done;
    int32_t match_count, has_resigned, can_resign0, can_resign1, resign0_set, resign1_set;
    float resign0_scores[GZP_MAX_ROLES], resign1_scores[GZP_MAX_ROLES];
    int32_t n_samples, starting_sample_depth, do_sample, in_run_to_end, run_to_end_can_resign;
    int32_t root_has_prev;
    int32_t error;
    int32_t mode;   // 0 self-play, 1 player (single search on command)
    // counters (monotonic)
    uint64_t total_evals, total_moves, total_games, total_samples, total_playouts;
    // traffic accounting for the roofline: nodes passed through by select, children
    // scanned there, children written by node creation
    uint64_t total_sel_nodes, total_sel_children, total_new_children;
    uint64_t total_transpositions;      // expansions that attached an existing node (stats.num_transpositions_attached)
    // per-manager style stats (selfplaymanager.h:46-72)
    int32_t saw_dupes, no_samples_taken, false_positive_resigns0, false_positive_resigns1,
        number_early_run_to_ends, number_resigns, number_aborts_game_length;
    int32_t error_restarts;  // games abandoned because a per-game capacity was exceeded (GZP_ERR_NODES / CHILDREN / PATH / SAMPLES)
    int32_t last_error;
    int32_t stop_request;    // player mode: the host's clock says stop (1: now, 2: as soon as the search has converged)
    int32_t start_pending;   // the next game of this slot starts from E->start_state[g] (bench / staggered starts)
};

// sample record header (int32 words), followed by state, prev state, policy pairs
enum {
    SR_READY = 0, SR_DEPTH, SR_GAME_LENGTH, SR_SLOT, SR_MATCH, SR_HAS_RESIGNED, SR_RESIGN_FP,
    SR_START_DEPTH, SR_VISITS, SR_LEAD, SR_N_PREV, SR_N_CHILDREN,
    SR_FINAL_SCORE = 12,      // float[4]
    SR_PUCT_SCORE = 16,       // float[4]
    SR_OTHER_MOVES = 20,      // int[4]: every role's action of child 0 (non-lead roles: their only move)
    SR_HEADER = 24
};

// One node: everything select / backup read and write, in one 64-byte line.
struct GZP_ALIGN(16) GzpNode {
    int32_t n_parent, n_child0, n_nchild;
    uint32_t n_visits;
    int32_t n_inflight, n_flags, n_lead, n_depth;
    float score[GZP_MAX_ROLES];        // current score per role
    float final_[GZP_MAX_ROLES];       // the network's (or the terminal) score per role
};
// One child of a node, 32 bytes: the edge (move, traversals, priors) AND a copy of the statistics of the node it
// leads to that the parent's select reads — that node's current score for the PARENT's lead role, its visit
// count, its finalised flag and in-flight count. The copy is kept current by whoever changes the child node:
// node creation (gzp_step_post), the network's answer (gzp_reply) and gzp_backup, which walks the playout path
// from the leaf up and therefore knows, at every node it updates, the parent's record that points to it.
// select then reads ONE contiguous block per tree level instead of gathering from every child node.
struct GZP_ALIGN(16) GzpChild {
    int32_t c_to;                       // node index or -1 (unexpanded)
    int16_t c_move;                     // the lead role's action index
    uint8_t c_flags, c_inflight;        // copy of the child node's NF_FINALISED bit / in-flight count
    uint32_t c_trav;
    float c_prob, c_prob_orig, c_next;
    float c_score;                      // copy of the child node's score for this node's lead role (-1: unexpanded)
    uint32_t c_visits;                  // copy of the child node's visit count
};

struct PuctEngine {
    int G, R, SW, LW, NCAP, CCAP, MAXD, n_rewards, max_actions;
    int role_sizes[GZP_MAX_ROLES];
    int legal_word_off[GZP_MAX_ROLES + 1];
    int policy_off[GZP_MAX_ROLES];      // offset of role r's policy row inside a concatenated row
    int policy_stride;                  // sum of role sizes
    // node records: index = g*NCAP + n (nodes of a game are in creation order: parent < child)
    GzpNode* nodes;                     // one 64-byte record per node
    uint32_t* n_state;                  // [.. * SW]
    int32_t* n_other;                   // [.. * GZP_MAX_ROLES] every role's only action (non-lead roles)
    // child records: index = g*CCAP + c (blocks in the creation order of their nodes)
    GzpChild* children;                 // one 32-byte record per child
    int32_t* c_owner;                   // node the child record belongs to (node creation and re-rooting only)
    int32_t *n_newidx, *n_newc0;        // re-root scratch: new node index / new child0
    // per game
    GameCtl* ctl;
    int32_t *path_node, *path_choice;   // [G*MAXD]
    uint32_t* root_prev_state;          // [G*SW] state of the root's parent (for samples)
    // pipeline scratch, one row per game
    uint32_t *x_in_state, *x_state, *x_legal;   // [G*SW], [G*SW], [G*LW]
    int16_t* x_moves;                           // [G*R]   (-1 = identity / inactive)
    uint8_t *x_term, *x_goals, *x_prevcnt;      // [G], [G*R], [G]
    uint8_t* x_need_eval;                       // [G*EB] 1 when the row of planes is a live request
    uint8_t* x_answered;                        // [G] or null: compacted batches — 1 when this step brought the game's
                                                // prediction (a request beyond the batch's cap waits one more step)
    // evaluation rows: EB per game (a Player with PuctConfig.batch_size > 1 hands up to batch_size leaves of its one
    // tree to the network at once; self-play: 1). Row of game g, request k: g*EB + k. The encoder reads e_state /
    // e_prev / e_prevcnt; with EB == 1 these ARE x_state / x_in_state / x_prevcnt.
    int EB;
    uint32_t *e_state, *e_prev;                 // [G*EB*SW]
    uint8_t* e_prevcnt;                         // [G*EB]
    GzpSched* sched;                            // [G] when EB > 1, else null
    int32_t *co_path_node, *co_path_choice;     // [G*EB*MAXD] playout path of every coroutine (EB > 1)
    uint32_t* x_cache;                          // [G * 12 * max_actions] child staging when it does not fit on chip
    // network outputs, one row per evaluation row
    const float* nn_policy;   // [G*EB * policy_stride]  role rows concatenated
    const float* nn_value;    // [G*EB * n_rewards]
    // samples
    int32_t* pend;            // [G * MAXPEND * srec] per-game pending samples
    int MAXPEND, srec;
    int32_t* ready;           // [READYCAP * srec] finished samples
    int READYCAP;
    unsigned long long* ready_count;   // [0] records appended since the last fetch, [1] records dropped (buffer full)
    // unique states: open-addressing table of 64-bit masked-state hashes with counters
    unsigned long long* uniq_keys;
    int32_t* uniq_counts;
    int uniq_cap;             // power of two
    const uint32_t* uniq_mask;  // [SW] createHashMask (gdltransformer.cpp:54-63)
    // transposition table (PuctConfig.lookup_transpositions, evaluator.cpp:144-173): per game an open-addressing
    // map masked state hash -> node index; tt_cap == 0: off
    unsigned long long* tt_keys;    // [G * tt_cap]
    int32_t* tt_node;               // [G * tt_cap]
    int tt_cap;                     // power of two >= 2 * NCAP
    const uint32_t* init_state; // [SW]
    uint32_t* start_state;      // [G*SW] per-slot start position of the slot's NEXT game when ctl.start_pending is set
    SpConf conf;
    uint64_t seed;
    // One engine step lets a game run until it needs the batch pipeline, but at most this many controller
    // iterations / tree levels descended: playouts that end in terminal or finalised nodes need no pipeline
    // pass, and the slowest warp sets the duration of the step for all games.
    int max_iters_per_step;
    int max_work_per_step;
    int prefetch_nodes;       // device select: prefetch the node records of a node's expanded children (gzp_warp_load_kid)
    long long* dbg_cycles;    // measurement aid (gz_puct_set_debug_cycles): per game [cycles in puct_pre_kernel, entry/exit phase, playouts, moves]
};

// ------------------------------------------------------------------------------------------------
// RNG: xoroshiro128+ (32-bit output), same mapping as oracle/shims/k273/rng.h
GZP_HD void gzp_rng_seed(uint64_t* st, uint64_t s) {
    for (int i = 0; i < 2; i++) {
        s += 0x9E3779B97F4A7C15ULL;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        st[i] = z ^ (z >> 31);
    }
}
#ifdef __CUDA_ARCH__
#define GZP_WARP_BARRIER() __syncwarp()
#else
#define GZP_WARP_BARRIER() ((void)0)
#endif
// On the device all 32 lanes run the controller redundantly against ONE copy of the generator state (shared
// memory): every lane must have read the old state before any lane writes the new one, whatever the hardware
// does about lockstep execution, hence the two warp barriers (called from warp-uniform control flow only).
GZP_HD uint32_t gzp_rng_next(uint64_t* st) {
    GZP_WARP_BARRIER();
    const uint64_t s0 = st[0];
    uint64_t s1 = st[1];
    GZP_WARP_BARRIER();
    const uint64_t r = s0 + s1;
    s1 ^= s0;
    st[0] = ((s0 << 55) | (s0 >> 9)) ^ s1 ^ (s1 << 14);
    st[1] = (s1 << 36) | (s1 >> 28);
    return (uint32_t)(r >> 32);
}
GZP_HD double gzp_rng_get(uint64_t* st) { return gzp_rng_next(st) * (1.0 / 4294967296.0); }
GZP_HD uint32_t gzp_rng_with_max(uint64_t* st, uint32_t n) { return n ? gzp_rng_next(st) % n : 0; }


// ------------------------------------------------------------------------------------------------
// Child ordering. The reference orders children with std::sort and comparators that leave many
// ties (and, in selectChild, are not even a strict weak order), so the visiting order under ties is
// whatever libstdc++'s introsort produces. To reproduce the reference's choices exactly the same
// algorithm is restated here (introsort: median-of-three to first, unguarded Hoare partition,
// recursion on the right part, depth limit 2*floor(log2 n) with heap-sort fallback, then the final
// insertion sort with threshold 16), operating on an array of child indices that starts in index
// order exactly like the reference's push_back loop (evaluator.cpp:244-248, node.cpp:318-322).
template <typename Cmp>
GZP_HDN void gzp_unguarded_linear_insert(int* a, int last, Cmp& cmp) {
    const int val = a[last];
    int next = last - 1;
    while (cmp(val, a[next])) { a[last] = a[next]; last = next; --next; }
    a[last] = val;
}
template <typename Cmp>
GZP_HDN void gzp_insertion_sort(int* a, int first, int last, Cmp& cmp) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (cmp(a[i], a[first])) {
            const int val = a[i];
            for (int k = i; k > first; --k) a[k] = a[k - 1];
            a[first] = val;
        } else {
            gzp_unguarded_linear_insert(a, i, cmp);
        }
    }
}
template <typename Cmp>
GZP_HDN void gzp_push_heap(int* a, int first, int hole, int top, int value, Cmp& cmp) {
    int parent = (hole - 1) / 2;
    while (hole > top && cmp(a[first + parent], value)) {
        a[first + hole] = a[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a[first + hole] = value;
}
template <typename Cmp>
GZP_HDN void gzp_adjust_heap(int* a, int first, int hole, int len, int value, Cmp& cmp) {
    const int top = hole;
    int second = hole;
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (cmp(a[first + second], a[first + (second - 1)])) second--;
        a[first + hole] = a[first + second];
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        a[first + hole] = a[first + (second - 1)];
        hole = second - 1;
    }
    gzp_push_heap(a, first, hole, top, value, cmp);
}
template <typename Cmp>
GZP_HDN void gzp_heap_sort(int* a, int first, int last, Cmp& cmp) {
    const int len = last - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            const int value = a[first + parent];
            gzp_adjust_heap(a, first, parent, len, value, cmp);
            if (parent == 0) break;
            parent--;
        }
    }
    int l = last;
    while (l - first > 1) {
        --l;
        const int value = a[l];
        a[l] = a[first];
        gzp_adjust_heap(a, first, 0, l - first, value, cmp);
    }
}
template <typename Cmp>
GZP_HDN void gzp_std_sort(int* a, int n, Cmp cmp) {
    if (n <= 0) return;
    // introsort loop, iterative: pending right-hand ranges on a small stack
    int stack_first[64], stack_last[64], stack_depth[64];
    int sp = 0;
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) lg++;
    stack_first[0] = 0; stack_last[0] = n; stack_depth[0] = 2 * lg; sp = 1;
    while (sp > 0) {
        --sp;
        int first = stack_first[sp], last = stack_last[sp], depth = stack_depth[sp];
        while (last - first > 16) {
            if (depth == 0) { gzp_heap_sort(a, first, last, cmp); break; }
            --depth;
            // median of (first+1, mid, last-1) moved to first
            const int mid = first + (last - first) / 2;
            const int ia = first + 1, ib = mid, ic = last - 1;
            int m;
            if (cmp(a[ia], a[ib])) {
                if (cmp(a[ib], a[ic])) m = ib;
                else if (cmp(a[ia], a[ic])) m = ic;
                else m = ia;
            } else if (cmp(a[ia], a[ic])) m = ia;
            else if (cmp(a[ib], a[ic])) m = ic;
            else m = ib;
            { const int t = a[first]; a[first] = a[m]; a[m] = t; }
            // unguarded partition of [first+1, last) around a[first]
            int lo = first + 1, hi = last;
            while (true) {
                while (cmp(a[lo], a[first])) ++lo;
                --hi;
                while (cmp(a[first], a[hi])) --hi;
                if (!(lo < hi)) break;
                const int t = a[lo]; a[lo] = a[hi]; a[hi] = t;
                ++lo;
            }
            const int cut = lo;
            if (sp < 64) { stack_first[sp] = cut; stack_last[sp] = last; stack_depth[sp] = depth; sp++; }
            last = cut;
        }
    }
    // final insertion sort
    if (n > 16) {
        gzp_insertion_sort(a, 0, 16, cmp);
        for (int i = 16; i != n; ++i) gzp_unguarded_linear_insert(a, i, cmp);
    } else {
        gzp_insertion_sort(a, 0, n, cmp);
    }
}

// ------------------------------------------------------------------------------------------------
// Staging area for the children of ONE node: the child records and the statistics of the nodes they
// point to. On the device it lives in shared memory (per warp) and is filled cooperatively by the 32
// lanes, so that the sequential decision logic below reads on-chip data instead of chasing pointers
// through HBM; on the host build it is plain memory filled by a loop.
struct ChildCache {
    int cap;
    int32_t *to, *move, *flags, *inflight, *order;
    uint32_t *trav, *visits;
    float *prob, *prob_orig, *next, *score, *tmp;
    double* expl;    // per-child exploration term of the current select (computed by the lanes in parallel)
};

GZP_HD size_t gzp_cache_words(int cap) { return (size_t)14 * (size_t)cap + 2; }

GZP_HD void gzp_cache_bind(ChildCache& C, uint32_t* mem, int cap) {
    C.cap = cap;
    C.to = (int32_t*)mem; mem += cap;
    C.move = (int32_t*)mem; mem += cap;
    C.flags = (int32_t*)mem; mem += cap;
    C.inflight = (int32_t*)mem; mem += cap;
    C.order = (int32_t*)mem; mem += cap;
    C.trav = mem; mem += cap;
    C.visits = mem; mem += cap;
    C.prob = (float*)mem; mem += cap;
    C.prob_orig = (float*)mem; mem += cap;
    C.next = (float*)mem; mem += cap;
    C.score = (float*)mem; mem += cap;
    C.tmp = (float*)mem; mem += cap;
    mem += ((uintptr_t)mem & 4) ? 1 : 0;          // 8-byte alignment for the doubles
    C.expl = (double*)mem;
}

struct GameView {
    const PuctEngine* E;
    int g;
    GameCtl* c;
    int64_t nb, cb;   // node / child base index of the active arena
    int lane;         // 0..31 on the device (all lanes run the controller redundantly), 0 on the host
    int32_t *pn, *pc; // playout path (node, chosen child); global rows, or a shared-memory copy on the device
    int32_t* ord;     // device: 32 words of shared memory per warp (rank -> lane permutation of the warp path)
    int work;         // tree levels descended in this step (the step's work budget, see gzp_step_pre)
    ChildCache C;
};

#ifdef __CUDA_ARCH__
#define GZP_LANES 32
#define GZP_SYNC() __syncwarp()
#else
#define GZP_LANES 1
#define GZP_SYNC() ((void)0)
#endif

GZP_HD GameView gzp_view(const PuctEngine* E, int g, int lane = 0, uint32_t* cache_mem = nullptr) {
    GameView v;
    v.E = E; v.g = g; v.c = &E->ctl[g];
    v.nb = (int64_t)g * E->NCAP;
    v.cb = (int64_t)g * E->CCAP;
    v.lane = lane;
    v.ord = nullptr;
    v.work = 0;
    v.pn = E->path_node + (int64_t)g * E->MAXD;
    v.pc = E->path_choice + (int64_t)g * E->MAXD;
    if (!cache_mem) cache_mem = E->x_cache + (int64_t)g * gzp_cache_words(E->max_actions);
    gzp_cache_bind(v.C, cache_mem, E->max_actions);
    return v;
}

#define NODE(field, n) (V.E->nodes[V.nb + (n)].field)
#define CHILD(field, c) (V.E->children[V.cb + (c)].field)
#define NSCORE(n, r) (V.E->nodes[V.nb + (n)].score[(r)])
#define NFINAL(n, r) (V.E->nodes[V.nb + (n)].final_[(r)])
#define NOTHER(n, r) (V.E->n_other[(V.nb + (n)) * GZP_MAX_ROLES + (r)])
#define NSTATE(n) (V.E->n_state + (V.nb + (n)) * V.E->SW)

GZP_HD const PuctConf& gzp_conf(const GameView& V) { return V.c->conf_sel ? V.E->conf.run_to_end : V.E->conf.puct; }
GZP_HD bool gzp_is_terminal(const GameView& V, int n) {
    return (NODE(n_flags, n) & NF_FORCE_TERMINAL) || NODE(n_nchild, n) == 0;
}
GZP_HD bool gzp_is_finalised(const GameView& V, int n) { return (NODE(n_flags, n) & NF_FINALISED) != 0; }

// Select must read the child nodes themselves (not the copies in the parent's child records) when a node can be
// reached through several records: transposition tables, and intra-tree workers (a child's node is attached to its
// record only when its creator resumes, and the reference's workers can create two nodes for one forced child).
GZP_HD bool gzp_direct_stats(const PuctEngine* E) { return E->tt_cap != 0 || E->EB > 1; }

// Intra-tree workers: how many coroutines are inside expandChild on a child of `parent` (child index `child`, or any
// child when child < 0) — PuctNodeChild::unselectable / PuctNode::unselectable_count (evaluator.cpp:228-233).
enum { CF_UNSELECTABLE = 0x100, CF_BLOCKED = 0x200 };    // bits of ChildCache::flags beside NF_FINALISED
GZP_HD int gzp_pending_on(const PuctEngine* E, int g, int parent, int child) {
    if (E->EB <= 1) return 0;
    const GzpSched& S = E->sched[g];
    if (!S.active) return 0;
    int n = 0;
    for (int k = 0; k <= E->EB; k++) {
        const GzpCo& co = S.co[k];
        if ((co.st == CO_AFTER_PROPNET || co.st == CO_AFTER_EVAL) && co.req_parent == parent &&
            (child < 0 || co.req_child == child)) n++;
    }
    return n;
}

// Stage node n's children in the cache: one contiguous block of records, two 16-byte loads per child.
GZP_RARE void gzp_cache_fill(GameView& V, int n) {
    const int nc = NODE(n_nchild, n), c0 = NODE(n_child0, n);
    ChildCache& C = V.C;
    GZP_SYNC();
    for (int i = V.lane; i < nc; i += GZP_LANES) {
        const GzpChild k = V.E->children[V.cb + c0 + i];
        C.to[i] = k.c_to;
        C.move[i] = k.c_move;
        C.trav[i] = k.c_trav;
        C.prob[i] = k.c_prob;
        C.prob_orig[i] = k.c_prob_orig;
        C.next[i] = k.c_next;
        C.score[i] = k.c_to >= 0 ? k.c_score : -1.0f;
        C.visits[i] = k.c_to >= 0 ? k.c_visits : 0u;
        C.flags[i] = k.c_to >= 0 ? (int)k.c_flags : 0;
        C.inflight[i] = k.c_to >= 0 ? (int)k.c_inflight : 0;
        if (gzp_direct_stats(V.E) && k.c_to >= 0) {   // the record's copy may be stale, read the node
            const int lead = NODE(n_lead, n);
            const GzpNode& t = V.E->nodes[V.nb + k.c_to];
            C.score[i] = t.score[lead < 0 ? 0 : lead];
            C.visits[i] = t.n_visits;
            C.flags[i] = t.n_flags & NF_FINALISED;
            C.inflight[i] = t.n_inflight;
            // every child of that node is being expanded by someone (evaluator.cpp:389-393)
            if (V.E->EB > 1 && t.n_nchild > 0 && gzp_pending_on(V.E, V.g, k.c_to, -1) == t.n_nchild) C.flags[i] |= CF_BLOCKED;
        }
        if (V.E->EB > 1 && gzp_pending_on(V.E, V.g, n, i) > 0) C.flags[i] |= CF_UNSELECTABLE;
    }
    GZP_SYNC();
}

// The parent's copy of node `to`'s statistics (see GzpChild): called wherever `to` changed and the record that
// points to it is known.
GZP_HD void gzp_child_refresh(GameView& V, int parent, int child, int to) {
    GzpChild& k = V.E->children[V.cb + NODE(n_child0, parent) + child];
    const int lead = NODE(n_lead, parent);
    k.c_score = NSCORE(to, lead < 0 ? 0 : lead);
    k.c_visits = NODE(n_visits, to);
    k.c_flags = (uint8_t)(NODE(n_flags, to) & NF_FINALISED);
    k.c_inflight = (uint8_t)(NODE(n_inflight, to) > 255 ? 255 : NODE(n_inflight, to));
}

// ------------------------------------------------------------------------------------------------
// Node creation from the propnet outputs of this game's scratch row (PuctNode::create + createNode).
// Returns the node index or -1 (arena full -> error flagged).
GZP_HDN int gzp_create_node(GameView& V, int parent, int depth) {
    const PuctEngine* E = V.E;
    GameCtl* c = V.c;
    const int g = V.g;
    const int R = E->R;
    if (c->n_nodes >= E->NCAP) { c->error = GZP_ERR_NODES; return -1; }
    const int n = c->n_nodes++;
    const uint32_t* legal = E->x_legal + (int64_t)g * E->LW;
    const bool terminal = E->x_term[g] != 0;
    int lead = 0, total_children = 0;
    int counts[GZP_MAX_ROLES];
    if (!terminal) {
        total_children = 1;
        int max_moves = 1;
        for (int r = 0; r < R; r++) {
            int cnt = 0;
            for (int w = E->legal_word_off[r]; w < E->legal_word_off[r + 1]; w++) {
#ifdef __CUDA_ARCH__
                cnt += __popc(legal[w]);
#else
                cnt += __builtin_popcount(legal[w]);
#endif
            }
            counts[r] = cnt;
            total_children *= cnt;
            if (cnt > max_moves) { max_moves = cnt; lead = r; }
        }
        if (max_moves > 1) {
            for (int r = 0; r < R; r++)
                if (r != lead && counts[r] > 1) { c->error = GZP_ERR_SIMULTANEOUS; lead = -1; }
        }
        if (lead < 0) return -1;
    }
    NODE(n_parent, n) = parent;
    NODE(n_visits, n) = 0;
    NODE(n_inflight, n) = 0;
    NODE(n_nchild, n) = total_children;
    NODE(n_flags, n) = terminal ? NF_FINALISED : 0;
    NODE(n_lead, n) = lead;
    NODE(n_depth, n) = depth;
    NODE(n_child0, n) = -1;
    for (int r = 0; r < GZP_MAX_ROLES; r++) { NSCORE(n, r) = 0.0f; NFINAL(n, r) = 0.0f; NOTHER(n, r) = 0; }
    uint32_t* st = NSTATE(n);
    const uint32_t* xs = E->x_state + (int64_t)g * E->SW;
    for (int w = 0; w < E->SW; w++) st[w] = xs[w];
    if (!terminal) {
        if (c->n_children + total_children > E->CCAP) { c->error = GZP_ERR_CHILDREN; c->n_nodes--; return -1; }
        const int c0 = c->n_children;
        c->n_children += total_children;
        NODE(n_child0, n) = c0;
        // every role's first legal action (the only one for non-lead roles)
        for (int r = 0; r < R; r++) {
            int first = 0;
            for (int w = E->legal_word_off[r]; w < E->legal_word_off[r + 1]; w++) {
                if (legal[w]) {
#ifdef __CUDA_ARCH__
                    first = (w - E->legal_word_off[r]) * 32 + (__ffs(legal[w]) - 1);
#else
                    first = (w - E->legal_word_off[r]) * 32 + __builtin_ctz(legal[w]);
#endif
                    break;
                }
            }
            NOTHER(n, r) = first;
        }
        // children = the lead role's legal actions in ascending index order
        int k = 0;
        for (int w = E->legal_word_off[lead]; w < E->legal_word_off[lead + 1]; w++) {
            uint32_t x = legal[w];
            while (x) {
#ifdef __CUDA_ARCH__
                const int b = __ffs(x) - 1;
#else
                const int b = __builtin_ctz(x);
#endif
                x &= x - 1;
                const int ci = c0 + k++;
                GzpChild k;
                k.c_to = -1;
                k.c_move = (int16_t)((w - E->legal_word_off[lead]) * 32 + b);
                k.c_flags = 0; k.c_inflight = 0;
                k.c_trav = 0;
                k.c_prob = 1.0f; k.c_prob_orig = 1.0f; k.c_next = 0.0f;
                k.c_score = -1.0f; k.c_visits = 0;
                V.E->children[V.cb + ci] = k;
                V.E->c_owner[V.cb + ci] = n;
            }
        }
    } else {
        for (int r = 0; r < R; r++) {
            const int gv = E->x_goals[(int64_t)g * R + r];
            const float s = (float)((gv == 255 ? 0 : gv) / 100.0);
            NFINAL(n, r) = s;
            float cs = s;
            // terminal score sharpening (evaluator.cpp:188-198)
            if (cs > 0.99) cs = (float)(cs * 1.05);
            else if (cs < 0.01) cs = -0.05f;
            NSCORE(n, r) = cs;
        }
    }
    return n;
}

// PuctNodeRequest::reply (node.cpp:463-511) from this game's row of network outputs.
GZP_HDN void gzp_reply(GameView& V, int n, int k = 0, bool refresh_parent = true) {
    const PuctEngine* E = V.E;
    const int64_t row = (int64_t)V.g * E->EB + k;        // evaluation row k of this game
    const int lead = NODE(n_lead, n);
    const int nc = NODE(n_nchild, n);
    const int c0 = NODE(n_child0, n);
    const float* pol = E->nn_policy + row * E->policy_stride + E->policy_off[lead];
    float* stage = V.C.tmp;
#ifdef __CUDA_ARCH__
    if (nc <= 32) {
        // warp path: child i in lane i; the sum runs in child order like the reference's loop
        GzpChild* kids = V.E->children + V.cb + c0;
        const bool act = V.lane < nc;
        float st = 0.0f;
        if (act) {
            const float p = pol[kids[V.lane].c_move];
            st = p > 0.001f ? p : 0.001f;
        }
        float total = 0.0f;
        for (int j = 0; j < nc; j++) total += __shfl_sync(0xffffffffu, st, j);
        if (act) {
            const float p = st / total;
            kids[V.lane].c_prob_orig = p;
            kids[V.lane].c_prob = p;
        }
        __syncwarp();
    } else
#endif
    {
    // gather the raw policy of every child (lanes in parallel), then sum in child order
    GZP_SYNC();
    for (int i = V.lane; i < nc; i += GZP_LANES) {
        const float p = pol[CHILD(c_move, c0 + i)];
        stage[i] = p > 0.001f ? p : 0.001f;          // std::max(0.001f, p)
    }
    GZP_SYNC();
    float total = 0.0f;
    for (int i = 0; i < nc; i++) total += stage[i];
    for (int i = V.lane; i < nc; i += GZP_LANES) {
        const float p = stage[i] / total;
        CHILD(c_prob_orig, c0 + i) = p;
        CHILD(c_prob, c0 + i) = p;
    }
    GZP_SYNC();
    }
    const float* val = E->nn_value + row * E->n_rewards;
    for (int r = 0; r < E->R; r++) {
        float s = val[r];
        if (E->n_rewards == 3) {
            const float mid = val[2] / 2.0f;
            s += mid;
        }
        if (s > 1.0) s = 1.0f;
        else if (s < 0.0) s = 0.0f;
        NFINAL(n, r) = s;
        NSCORE(n, r) = s;   // clamp of an already clamped value
    }
    // n is the node the game's last request created: its parent's record keeps a copy of its score
    if (refresh_parent && V.c->req_parent >= 0) gzp_child_refresh(V, V.c->req_parent, V.c->req_child, n);
}


// Orders the staged children into C.order exactly as the reference's std::sort call would. For
// n <= 16 libstdc++'s std::sort is a plain (stable) insertion sort and all comparators used here are
// strict weak orders, so on the device every lane computes the final rank of one child directly;
// larger nodes go through the restated introsort.
template <typename Cmp>
GZP_HDN void gzp_order_children(GameView& V, int nc, Cmp cmp) {
    int* order = V.C.order;
#ifdef __CUDA_ARCH__
    // Every lane ranks its children directly: rank(i) = #{j before i} under the stable order.
    //  * n <= 16: std::sort IS a stable insertion sort, so this is exact.
    //  * n  > 16: introsort is not stable, but when no two children compare equivalent the sorted
    //    order is unique and any correct sort agrees with it; only in the presence of ties the
    //    restated introsort below decides (sequentially, as the reference's library would).
    __syncwarp();
    bool tie = false;
    for (int i = V.lane; i < nc; i += 32) {
        int rank = 0;
        for (int j = 0; j < nc; j++) {
            if (j == i) continue;
            const bool ji = cmp(j, i), ij = cmp(i, j);
            if (ji || (!ij && j < i)) rank++;
            if (!ji && !ij) tie = true;
        }
        reinterpret_cast<int*>(V.C.tmp)[i] = rank;
    }
    const bool any_tie = __any_sync(0xffffffffu, tie);
    if (nc <= 16 || !any_tie) {
        for (int i = V.lane; i < nc; i += 32) order[reinterpret_cast<int*>(V.C.tmp)[i]] = i;
        __syncwarp();
        return;
    }
    __syncwarp();
#endif
    for (int i = 0; i < nc; i++) order[i] = i;
    gzp_std_sort(order, nc, cmp);
}

#ifdef __CUDA_ARCH__
// Wide nodes (hexLG13 170 children, connect6 361, amazons 2176): ranking every child against every other through the
// comparator costs two comparator calls per pair. Both orders used on the select path are lexicographic orders
// of plain values, so a child gets ONE 64-bit key (larger = earlier) and the ranking loop is a load, a compare and
// an add per pair. Equal keys (= children the comparator calls equivalent) hand the node back to
// gzp_order_children, whose restated std::sort decides ties as the reference's library does.
__device__ __forceinline__ unsigned gzp_fkey(float f) {          // monotonic float -> unsigned (-0 -> +0 first)
    const unsigned u = __float_as_uint(f + 0.0f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __noinline__ bool gzp_order_by_key(GameView& V, int nc, const unsigned long long* key) {
    int* order = V.C.order;
    int* rk = reinterpret_cast<int*>(V.C.tmp);
    __syncwarp();
    bool tie = false;
    for (int i = V.lane; i < nc; i += 32) {
        const unsigned long long ki = key[i];
        int rank = 0;
        for (int j = 0; j < nc; j++) {
            const unsigned long long kj = key[j];
            rank += (kj > ki || (kj == ki && j < i)) ? 1 : 0;
            tie |= kj == ki && j != i;
        }
        rk[i] = rank;
    }
    if (__any_sync(0xffffffffu, tie)) {
        // equivalent children (a synthetic network floors many priors to the same 0.001): the restated introsort
        // decides, as the reference's library would — with the keys as its comparator (the same truth table as
        // the value comparator for every pair, a third of its instructions)
        __syncwarp();
        for (int i = 0; i < nc; i++) order[i] = i;
        gzp_std_sort(order, nc, [key](int a, int b) { return key[a] > key[b]; });
        __syncwarp();
        return true;
    }
    for (int i = V.lane; i < nc; i += 32) order[rk[i]] = i;
    __syncwarp();
    return true;
}
#endif

// ------------------------------------------------------------------------------------------------
// chooseTopVisits (evaluator.cpp:1100-1159). Sort key: traversals desc, then policy_prob desc
// (node.cpp:346-373), ordered with the restated std::sort.
// Expects the cache to hold node n's children (gzp_cache_fill).
GZP_RARE int gzp_choose_top_visits(const GameView& V, int n, int nc_known = -1) {
    const int nc = nc_known >= 0 ? nc_known : NODE(n_nchild, n);
    const PuctConf& cf = gzp_conf(V);
    const ChildCache& C = V.C;
    int* order = C.order;
    bool ordered = false;
#ifdef __CUDA_ARCH__
    if (nc > 32) {
        unsigned long long* key = reinterpret_cast<unsigned long long*>(C.expl);     // (free until select fills it)
        __syncwarp();
        for (int i = V.lane; i < nc; i += 32)
            key[i] = ((unsigned long long)(unsigned)(int)C.trav[i] << 32) | gzp_fkey(C.prob[i]);
        ordered = gzp_order_by_key(const_cast<GameView&>(V), nc, key);
    }
#endif
    if (!ordered)
    gzp_order_children(const_cast<GameView&>(V), nc, [&C](int a, int b) {
        const int ta = (int)C.trav[a], tb = (int)C.trav[b];
        if (ta == tb) return C.prob[a] > C.prob[b];
        return ta > tb;
    });
    int indx0 = -1, indx1 = -1;
    int count = 0;
    for (int k = 0; k < nc; k++) {
        const int ci = order[k];
        if (C.to[ci] >= 0 && (C.flags[ci] & NF_FINALISED)) {
            const float sc = C.score[ci];
            if (sc > 0.99) return ci;
            if (sc < 0.01) { count++; continue; }
        }
        if (indx0 == -1) indx0 = count;
        else if (indx1 == -1) indx1 = count;
        count++;
    }
    if (cf.top_visits_best_guess_converge_ratio > 0 && indx0 != -1 && indx1 != -1) {
        const int i0 = order[indx0], i1 = order[indx1];
        if (C.to[i0] >= 0 && C.to[i1] >= 0) {
            if ((float)C.trav[i1] > (float)C.trav[i0] * cf.top_visits_best_guess_converge_ratio &&
                C.score[i1] > C.score[i0])
                return i1;
            return i0;
        }
    }
    return order[0];
}

// priorScore (evaluator.cpp:1195-1224)
// node_final = the node's final score for its lead role, node_visits / nc its visit and child counts — the
// caller has them in registers already (each would otherwise be a dependent global load on the descent).
GZP_RARE float gzp_prior_score(const GameView& V, int n, int depth, float node_final, uint32_t node_visits, int nc) {
    const PuctConf& cf = gzp_conf(V);
    const ChildCache& C = V.C;
    float prior = node_final;
    if (node_visits > 8) {
        const int best = gzp_choose_top_visits(V, n, nc);
        if (C.to[best] >= 0) prior = C.score[best];
    }
    float fpu = depth == 0 ? cf.fpu_prior_discount_root : cf.fpu_prior_discount;
    if (fpu > 0) {
        float total = 0.0f;
        for (int i = 0; i < nc; i++)
            if (C.to[i] >= 0 && C.visits[i] > 0) total += C.prob[i];
        fpu *= sqrtf(total);
        prior -= fpu;
    }
    return prior;
}

// The evaluator's generator as a 32-bit URBG for the restated library samplers (gz_libm.h).
struct GzpRng32 {
    uint64_t* st;
    GZP_MEMBER uint32_t operator()() { return gzp_rng_next(st); }
};

// setDirichletNoise (evaluator.cpp:1227-1297)
GZP_RARE void gzp_set_dirichlet_noise(GameView& V, int n) {
    float* scratch = V.C.tmp;
    const PuctConf& cf = gzp_conf(V);
    const int nc = NODE(n_nchild, n), c0 = NODE(n_child0, n);
    if ((NODE(n_flags, n) & NF_NOISE_SET) || nc < 2 || cf.dirichlet_noise_pct < 0) return;
    if (NSCORE(n, NODE(n_lead, n)) > 0.95) return;
    // std::gamma_distribution<float> gamma(alpha, 1.0f), one object for all children (evaluator.cpp:1249-1258):
    // restated bit for bit in gz_libm.h, so the noise vector is the reference's
    const float alpha = 10.83f / nc;
    GzGammaF gamma;
    gz_gamma_init(gamma, alpha);
    GzpRng32 rng{V.c->rng_eval};
    float total_noise = 0.0f;
    for (int i = 0; i < nc; i++) {
        const float x = gz_gamma_float(gamma, rng);
        scratch[i] = x;
        total_noise += x;
    }
    if (total_noise < 1.17549435e-38f) return;
    for (int i = 0; i < nc; i++) scratch[i] /= total_noise;
    const bool squash = cf.noise_policy_squash_pct > 0 && gzp_rng_get(V.c->rng_eval) < cf.noise_policy_squash_pct;
    float total_policy = 0;
    const float pct = cf.dirichlet_noise_pct;
    for (int i = 0; i < nc; i++) {
        float p = CHILD(c_prob, c0 + i);
        if (squash) p = cf.noise_policy_squash_prob < p ? cf.noise_policy_squash_prob : p;
        p = (1.0f - pct) * p + pct * scratch[i];
        CHILD(c_prob, c0 + i) = p;
        total_policy += p;
    }
    for (int i = 0; i < nc; i++) {
        const float p = CHILD(c_prob, c0 + i) / total_policy;
        CHILD(c_prob, c0 + i) = p;
        V.C.prob[i] = p;     // keep the staged copy current
    }
    NODE(n_flags, n) |= NF_NOISE_SET;
}

struct GzpKid {          // warp path: one lane's child, unpacked (lanes >= nc: an unexpanded dummy)
    int to, flags, inflight;
    uint32_t trav, visits;
    float prob, prob_orig, score;
    bool act, expanded, fin;
};

#ifdef __CUDA_ARCH__
// ------------------------------------------------------------------------------------------------
// Warp path (device only): nodes with at most 32 children keep child i in lane i's REGISTERS — the 32-byte
// record is two 16-byte loads per lane, one coalesced block per tree level — and every per-child quantity of
// selectChild / priorScore / chooseTopVisits / converged is computed by the child's own lane. None of them
// needs the reference's sorted order as such, only its first element(s) with some property, which is a
// max-reduction over (key, key, index) as long as the maximum is unique — and when it is not, libstdc++'s
// insertion sort (n <= 16) keeps equal elements in index order, so the lowest lane wins. Results are those of
// the generic code below, which still handles everything unusual: more than 32 children, ties among more
// than 16 children (libstdc++'s introsort decides those), the root latch limiter, children with evaluations
// in flight, nodes whose children are all proven losses, and the node that receives Dirichlet noise. A
// warp-path function returns GZP_WARP_PUNT in those cases and the caller falls through to the generic code.
#define GZP_WARP_PUNT (-2)
#define GZP_FULL 0xffffffffu

__device__ __forceinline__ unsigned gzp_ordered_bits(float f) {      // monotonic float -> unsigned
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// The first element of the reference's sort by (k1 desc, k2 desc) among the lanes with `pred`:
// the lane, -1 when no lane has `pred`, GZP_WARP_PUNT when the maximum is shared and n > 16.
__device__ __forceinline__ int gzp_warp_top(bool pred, unsigned k1, float k2, int nc) {
    const unsigned any = __ballot_sync(GZP_FULL, pred);
    if (!any) return -1;
    const unsigned m1 = __reduce_max_sync(GZP_FULL, pred ? k1 : 0u);
    const bool p1 = pred && k1 == m1;
    const unsigned m2 = __reduce_max_sync(GZP_FULL, p1 ? gzp_ordered_bits(k2) : 0u);
    const unsigned top = __ballot_sync(GZP_FULL, p1 && gzp_ordered_bits(k2) == m2);
    if (__popc(top) > 1 && nc > 16) return GZP_WARP_PUNT;
    return __ffs(top) - 1;
}

__device__ __forceinline__ GzpKid gzp_warp_load_kid(const GameView& V, int c0, int nc, int lead) {
    GzpKid K;
    K.act = V.lane < nc;
    GzpChild k;
    k.c_to = -1; k.c_move = 0; k.c_flags = 0; k.c_inflight = 0; k.c_trav = 0;
    k.c_prob = 0.0f; k.c_prob_orig = 0.0f; k.c_next = 0.0f; k.c_score = -1.0f; k.c_visits = 0;
    if (K.act) k = V.E->children[V.cb + c0 + V.lane];
    K.to = k.c_to;
    K.expanded = K.act && k.c_to >= 0;
    K.trav = k.c_trav;
    K.prob = k.c_prob;
    K.prob_orig = k.c_prob_orig;
    K.score = K.expanded ? k.c_score : -1.0f;
    K.visits = K.expanded ? k.c_visits : 0u;
    K.flags = K.expanded ? (int)k.c_flags : 0;
    K.inflight = K.expanded ? (int)k.c_inflight : 0;
    // (measurement switch GZ_PUCT_PREFETCH=1) Whichever child select picks, the next tree level starts with that
    // child's node record: every lane asks L2 for its own child's record now, so that the DRAM round trip runs
    // under the select arithmetic instead of after it. Measured: no gain, see gz_puct.cu.
    if (K.expanded && V.E->prefetch_nodes)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(V.E->nodes + V.nb + k.c_to));
    if (gzp_direct_stats(V.E) && K.expanded) {     // the record's copy may be stale, read the node
        const GzpNode& t = V.E->nodes[V.nb + k.c_to];
        K.score = t.score[lead < 0 ? 0 : lead];
        K.visits = t.n_visits;
        K.flags = t.n_flags & NF_FINALISED;
        K.inflight = t.n_inflight;
    }
    K.fin = K.expanded && (K.flags & NF_FINALISED) != 0;
    return K;
}

// chooseTopVisits (evaluator.cpp:1100-1159): the sorted order is (traversals desc, prior desc)
__device__ __forceinline__ int gzp_warp_top_visits(const GameView& V, const GzpKid& K, int nc, float ratio) {
    const bool win = K.fin && K.score > 0.99;
    const bool loss = K.fin && K.score < 0.01;
    const int w = gzp_warp_top(K.act && win, K.trav, K.prob, nc);        // the first proven win in sorted order
    if (w != -1) return w;
    const bool cand = K.act && !loss;
    const int i0 = gzp_warp_top(cand, K.trav, K.prob, nc);
    if (i0 == GZP_WARP_PUNT) return GZP_WARP_PUNT;
    if (ratio > 0 && i0 >= 0) {
        const int i1 = gzp_warp_top(cand && V.lane != i0, K.trav, K.prob, nc);
        if (i1 == GZP_WARP_PUNT) return GZP_WARP_PUNT;
        if (i1 >= 0) {
            const int to0 = __shfl_sync(GZP_FULL, K.to, i0), to1 = __shfl_sync(GZP_FULL, K.to, i1);
            if (to0 >= 0 && to1 >= 0) {
                const float t0 = (float)__shfl_sync(GZP_FULL, K.trav, i0), t1 = (float)__shfl_sync(GZP_FULL, K.trav, i1);
                const float s0 = __shfl_sync(GZP_FULL, K.score, i0), s1 = __shfl_sync(GZP_FULL, K.score, i1);
                if (t1 > t0 * ratio && s1 > s0) return i1;
                return i0;
            }
        }
    }
    return gzp_warp_top(K.act, K.trav, K.prob, nc);                       // order[0]
}

// converged (evaluator.cpp:1342-1362): the two first children by (node visits desc, prior desc)
__device__ __forceinline__ int gzp_warp_converged(const GameView& V, const GzpKid& K, int nc, int count) {
    const int a = gzp_warp_top(K.act, K.visits, K.prob, nc);
    if (a == GZP_WARP_PUNT) return GZP_WARP_PUNT;
    const int b = gzp_warp_top(K.act && V.lane != a, K.visits, K.prob, nc);
    if (b == GZP_WARP_PUNT) return GZP_WARP_PUNT;
    const int toa = __shfl_sync(GZP_FULL, K.to, a), tob = __shfl_sync(GZP_FULL, K.to, b);
    if (toa >= 0 && tob >= 0) {
        const float sa = __shfl_sync(GZP_FULL, K.score, a), sb = __shfl_sync(GZP_FULL, K.score, b);
        const uint32_t va = __shfl_sync(GZP_FULL, K.visits, a), vb = __shfl_sync(GZP_FULL, K.visits, b);
        if (sa > sb && va > vb + (uint32_t)count) return 1;
    }
    return 0;
}

// selectChild (evaluator.cpp:341-517). The reference scans the children in sortedChildrenSelect order, keeps
// its running best score as a FLOAT and lets a child replace the best when its DOUBLE score exceeds it. With F
// the largest float-rounded score, the first child (in scan order) that rounds to F always replaces the best
// (rounding is monotonic), nothing that rounds below F can replace it afterwards, and a later child that
// rounds to F replaces it only if its double score is above F. So the scan order matters only among the
// children that round to F: usually one, and then no order is needed at all.
__device__ int gzp_warp_select(GameView& V, const GzpKid& K, int nc, uint32_t node_visits, float node_final,
                               float puct_c, int depth, int* chosen_to) {
    const PuctConf& cf = gzp_conf(V);
    const int lane = V.lane;
    if (node_visits > 1000 && node_visits < 40000000 && depth == 0) return GZP_WARP_PUNT;      // root latch limiter
    if (__any_sync(GZP_FULL, K.act && K.trav > 0 && K.inflight > 0)) return GZP_WARP_PUNT;     // virtual-loss discount
    // ---- priorScore (evaluator.cpp:1195-1224)
    float prior = node_final;
    if (node_visits > 8) {
        const int best = gzp_warp_top_visits(V, K, nc, cf.top_visits_best_guess_converge_ratio);
        if (best == GZP_WARP_PUNT) return GZP_WARP_PUNT;
        const int bto = __shfl_sync(GZP_FULL, K.to, best);
        const float bsc = __shfl_sync(GZP_FULL, K.score, best);
        if (bto >= 0) prior = bsc;
    }
    float fpu = depth == 0 ? cf.fpu_prior_discount_root : cf.fpu_prior_discount;
    if (fpu > 0) {
        const float mine = (K.expanded && K.visits > 0) ? K.prob : 0.0f;
        float total = 0.0f;
        for (int j = 0; j < nc; j++) total += __shfl_sync(GZP_FULL, mine, j);       // in child order, as the reference sums
        fpu *= sqrtf(total);
        prior -= fpu;
    }
    // ---- this lane's child: score + exploration (evaluator.cpp:393-457)
    const double sqrt_node_visits = sqrt((double)(node_visits + 1u));
    double exploration = (double)(puct_c * K.prob) * sqrt_node_visits / ((int)K.trav + 1 + (double)K.inflight);
    double child_score = prior;
    bool badfb = false, winret = false;
    if (K.expanded) {
        child_score = K.score;
        if (K.fin) {
            if (child_score > 0.99) {
                if (depth > 0) winret = true;
                else child_score *= 1.0f + puct_c;
            } else if (child_score < 0.01) {
                badfb = true;
            } else {
                exploration = 0.0;
            }
        }
    }
    int best_child;
    const unsigned wm = __ballot_sync(GZP_FULL, K.act && winret);
    if (wm) {
        // a proven win below the root is taken at once: the first one in scan order, i.e. the highest score
        // (expanded children sort by score); equal scores keep index order up to 16 children
        const unsigned m = __reduce_max_sync(GZP_FULL, (K.act && winret) ? gzp_ordered_bits(K.score) : 0u);
        const unsigned top = __ballot_sync(GZP_FULL, K.act && winret && gzp_ordered_bits(K.score) == m);
        if (__popc(top) > 1 && nc > 16) return GZP_WARP_PUNT;
        best_child = __ffs(top) - 1;
    } else {
        const bool elig = K.act && !badfb;
        const double score = child_score + exploration;
        const float f = (float)score;
        const unsigned Fu = __reduce_max_sync(GZP_FULL, elig ? gzp_ordered_bits(f) : 0u);
        const unsigned S = __ballot_sync(GZP_FULL, elig && gzp_ordered_bits(f) == Fu);
        if (!S) return GZP_WARP_PUNT;                     // only proven losses: the generic scan picks the fallback
        const float F = __shfl_sync(GZP_FULL, f, __ffs(S) - 1);
        if (!(F > -1.0f)) return GZP_WARP_PUNT;           // (float best_score = -1 is never reached by a real score)
        if (__popc(S) == 1) {
            best_child = __ffs(S) - 1;
        } else {
            if (nc > 16) return GZP_WARP_PUNT;
            // several children round to F: their scan order decides (stable order of sortedChildrenSelect's
            // comparator, evaluator.cpp:242-263). pos = this lane's position among them.
            int pos = 0;
            const bool inS = (S >> lane) & 1u;
            for (unsigned rest = S; rest; rest &= rest - 1) {
                const int j = __ffs(rest) - 1;
                const float sb = __shfl_sync(GZP_FULL, K.score, j), pob = __shfl_sync(GZP_FULL, K.prob_orig, j);
                const bool both = K.score < 0 && sb < 0;
                const bool ji = both ? (pob > K.prob_orig) : (sb > K.score);
                const bool ij = both ? (K.prob_orig > pob) : (K.score > sb);
                if (j != lane && (ji || (!ij && j < lane))) pos++;
            }
            // the first of them takes the lead; a later one replaces it only with a double score above F
            const unsigned later = __ballot_sync(GZP_FULL, inS && pos > 0 && score > (double)F);
            if (later) {
                const unsigned m = __reduce_max_sync(GZP_FULL, (inS && pos > 0 && score > (double)F) ? (unsigned)pos : 0u);
                best_child = __ffs(__ballot_sync(GZP_FULL, inS && (unsigned)pos == m)) - 1;
            } else {
                best_child = __ffs(__ballot_sync(GZP_FULL, inS && pos == 0)) - 1;
            }
        }
    }
    *chosen_to = __shfl_sync(GZP_FULL, K.to, best_child);
    return best_child;
}

// ------------------------------------------------------------------------------------------------
// Wide nodes on the device (more than 32 children: hexLG13 170, connect6 361, chess): the same order-free
// formulation as gzp_warp_select / gzp_warp_top_visits, over the children staged in shared memory, every lane
// taking children lane, lane+32, ... — no ordering pass, no sequential scan. Anything the formulation does not
// cover (the root latch limiter, evaluations in flight below the node, candidates that tie for the maximum) is
// handed to the generic path, which reproduces the reference's scan step by step.
__device__ __noinline__ int gzp_wide_top(const GameView& V, int nc, int which, int skip) {
    // which: 0 all children, 1 proven wins, 2 everything but proven losses; skip: a child to leave out (-1: none)
    // key: (traversals desc, prior desc) as chooseTopVisits orders (node.cpp:346-373)
    const ChildCache& C = V.C;
    unsigned m1 = 0, any = 0;
    for (int i = V.lane; i < nc; i += 32) {
        const bool fin = C.to[i] >= 0 && (C.flags[i] & NF_FINALISED);
        const bool win = fin && C.score[i] > 0.99, loss = fin && C.score[i] < 0.01;
        const bool pred = i != skip && (which == 0 || (which == 1 ? win : !loss));
        if (pred) { any = 1; m1 = C.trav[i] > m1 ? C.trav[i] : m1; }
    }
    if (!__any_sync(GZP_FULL, any != 0)) return -1;
    m1 = __reduce_max_sync(GZP_FULL, any ? m1 : 0u);
    unsigned m2 = 0;
    for (int i = V.lane; i < nc; i += 32) {
        const bool fin = C.to[i] >= 0 && (C.flags[i] & NF_FINALISED);
        const bool win = fin && C.score[i] > 0.99, loss = fin && C.score[i] < 0.01;
        const bool pred = i != skip && (which == 0 || (which == 1 ? win : !loss));
        if (pred && C.trav[i] == m1) { const unsigned k = gzp_ordered_bits(C.prob[i]); m2 = k > m2 ? k : m2; }
    }
    m2 = __reduce_max_sync(GZP_FULL, m2);
    int count = 0, idx = 0x7fffffff;
    for (int i = V.lane; i < nc; i += 32) {
        const bool fin = C.to[i] >= 0 && (C.flags[i] & NF_FINALISED);
        const bool win = fin && C.score[i] > 0.99, loss = fin && C.score[i] < 0.01;
        const bool pred = i != skip && (which == 0 || (which == 1 ? win : !loss));
        if (pred && C.trav[i] == m1 && gzp_ordered_bits(C.prob[i]) == m2) { count++; idx = i < idx ? i : idx; }
    }
    count = __reduce_add_sync(GZP_FULL, count);
    if (count > 1) return GZP_WARP_PUNT;          // equivalent candidates: std::sort's tie order decides (nc > 16 here)
    return __reduce_min_sync(GZP_FULL, idx);
}

__device__ __noinline__ int gzp_wide_top_visits(const GameView& V, int nc, float ratio) {
    const ChildCache& C = V.C;
    const int w = gzp_wide_top(V, nc, 1, -1);              // the first proven win in sorted order
    if (w != -1) return w;
    const int i0 = gzp_wide_top(V, nc, 2, -1);
    if (i0 == GZP_WARP_PUNT) return GZP_WARP_PUNT;
    if (ratio > 0 && i0 >= 0) {
        const int i1 = gzp_wide_top(V, nc, 2, i0);
        if (i1 == GZP_WARP_PUNT) return GZP_WARP_PUNT;
        if (i1 >= 0 && C.to[i0] >= 0 && C.to[i1] >= 0) {
            const float t0 = (float)C.trav[i0], t1 = (float)C.trav[i1];
            if (t1 > t0 * ratio && C.score[i1] > C.score[i0]) return i1;
            return i0;
        }
    }
    return gzp_wide_top(V, nc, 0, -1);                      // order[0]
}

__device__ __noinline__ int gzp_wide_select(GameView& V, int n, int nc, uint32_t node_visits, float node_final, float puct_c,
                                            int depth, int* chosen_to) {
    const PuctConf& cf = gzp_conf(V);
    ChildCache& C = V.C;
    if (node_visits > 1000 && node_visits < 40000000 && depth == 0) return GZP_WARP_PUNT;      // root latch limiter
    gzp_cache_fill(V, n);
    bool vl = false;
    for (int i = V.lane; i < nc; i += 32) vl |= C.trav[i] > 0 && C.to[i] >= 0 && C.inflight[i] > 0;
    if (__any_sync(GZP_FULL, vl)) return GZP_WARP_PUNT;                                       // virtual-loss discount
    // ---- priorScore (evaluator.cpp:1195-1224)
    float prior = node_final;
    if (node_visits > 8) {
        const int best = gzp_wide_top_visits(V, nc, cf.top_visits_best_guess_converge_ratio);
        if (best == GZP_WARP_PUNT) return GZP_WARP_PUNT;
        if (best >= 0 && C.to[best] >= 0) prior = C.score[best];
    }
    float fpu = depth == 0 ? cf.fpu_prior_discount_root : cf.fpu_prior_discount;
    if (fpu > 0) {
        float total = 0.0f;                      // in child order, as the reference sums
        for (int j = 0; j < nc; j++) total += (C.to[j] >= 0 && C.visits[j] > 0) ? C.prob[j] : 0.0f;
        fpu *= sqrtf(total);
        prior -= fpu;
    }
    // ---- every child's score + exploration (evaluator.cpp:393-457), float-rounded into C.tmp; -inf = not eligible
    const double sqrt_node_visits = sqrt((double)(node_visits + 1u));
    const float neg_inf = -3.402823466e+38f;
    unsigned win_best = 0;
    int win_cnt = 0, win_idx = 0x7fffffff;
    __syncwarp();
    for (int i = V.lane; i < nc; i += 32) {
        const bool expanded = C.to[i] >= 0;
        double exploration = (double)(puct_c * C.prob[i]) * sqrt_node_visits / ((int)C.trav[i] + 1 + (expanded ? (double)C.inflight[i] : 0.0));
        double child_score = prior;
        bool badfb = false, winret = false;
        if (expanded) {
            child_score = C.score[i];
            if (C.flags[i] & NF_FINALISED) {
                if (child_score > 0.99) {
                    if (depth > 0) winret = true;
                    else child_score *= 1.0f + puct_c;
                } else if (child_score < 0.01) {
                    badfb = true;
                } else {
                    exploration = 0.0;
                }
            }
        }
        C.tmp[i] = badfb ? neg_inf : (float)(child_score + exploration);
        if (winret) { const unsigned k = gzp_ordered_bits(C.score[i]); if (k > win_best) win_best = k; }
    }
    __syncwarp();
    int best_child;
    const unsigned wb = __reduce_max_sync(GZP_FULL, win_best);
    if (wb) {
        // a proven win below the root is taken at once: the highest score; equal scores -> the generic scan decides
        for (int i = V.lane; i < nc; i += 32) {
            const bool winret = depth > 0 && C.to[i] >= 0 && (C.flags[i] & NF_FINALISED) && C.score[i] > 0.99;
            if (winret && gzp_ordered_bits(C.score[i]) == wb) { win_cnt++; win_idx = i < win_idx ? i : win_idx; }
        }
        if (__reduce_add_sync(GZP_FULL, win_cnt) > 1) return GZP_WARP_PUNT;
        best_child = __reduce_min_sync(GZP_FULL, win_idx);
    } else {
        unsigned fu = 0;
        for (int i = V.lane; i < nc; i += 32)
            if (C.tmp[i] != neg_inf) { const unsigned k = gzp_ordered_bits(C.tmp[i]); fu = k > fu ? k : fu; }
        const unsigned Fu = __reduce_max_sync(GZP_FULL, fu);
        if (!Fu) return GZP_WARP_PUNT;                   // only proven losses: the generic scan picks the fallback
        int cnt = 0, idx = 0x7fffffff;
        float F = 0.0f;
        for (int i = V.lane; i < nc; i += 32)
            if (C.tmp[i] != neg_inf && gzp_ordered_bits(C.tmp[i]) == Fu) { cnt++; idx = i < idx ? i : idx; F = C.tmp[i]; }
        if (__reduce_add_sync(GZP_FULL, cnt) != 1) return GZP_WARP_PUNT;      // several children round to F: scan order decides
        best_child = __reduce_min_sync(GZP_FULL, idx);
        F = C.tmp[best_child];
        if (!(F > -1.0f)) return GZP_WARP_PUNT;          // (float best_score = -1 is never reached by a real score)
    }
    *chosen_to = C.to[best_child];
    return best_child;
}
#endif  // __CUDA_ARCH__

// selectChild (evaluator.cpp:341-517). Appends to the path; returns the child index within the
// node (or -1 when nothing is selectable, which cannot happen with one evaluation in flight).
struct GzpSelInfo {
    int to; int inflight;   // the chosen child's node (or -1) and node n's in-flight count, read with the header
    int yield_first;        // intra-tree workers: the reference yields before returning this child (evaluator.cpp:492-497)
    int blocked;            // ... nothing selectable, every candidate is being expanded (stats.num_blocked)
};

// The generic form of selectChild: children staged in the cache, the reference's sorted scan reproduced step by
// step (any child count, the root latch limiter and the virtual-loss discount with their random draws,
// Dirichlet noise). The warp path above covers the common case on the device.
GZP_RARE int gzp_select_child_generic(GameView& V, int n, const GzpNode& nd, GzpSelInfo* info, float puct_c) {
    GameCtl* c = V.c;
    const PuctConf& cf = gzp_conf(V);
    const int depth = c->path_len;
    const int nc = nd.n_nchild;
    const int lead = nd.n_lead;
    int32_t* pn = V.pn;
    int32_t* pc = V.pc;
    const uint32_t node_visits = nd.n_visits;
    float node_finals[GZP_MAX_ROLES];
#pragma unroll
    for (int r = 0; r < GZP_MAX_ROLES; r++) node_finals[r] = nd.final_[r];
    gzp_cache_fill(V, n);
    ChildCache& C = V.C;
    if (depth == 0) gzp_set_dirichlet_noise(V, n);

    const float prior_score = gzp_prior_score(V, n, depth, node_finals[lead < 0 ? 0 : lead], node_visits, nc);
    const double sqrt_node_visits = sqrt((double)(node_visits + 1u));

    float best_score = -1;
    int best_child = -1;
    float best_child_score_actual = -1;
    int best_child_score = -1;
    int bad_fallback = -1;
    float best_fallback_score = -1;
    int best_fallback = -1;
    int unselectables = 0;

    // iterate children in sortedChildrenSelect order (evaluator.cpp:242-263): key sa = node score
    // for the lead role (-1 when unexpanded); both negative -> policy_prob_orig desc, else sa desc.
    int* order = C.order;
    bool ordered = false;
#ifdef __CUDA_ARCH__
    if (nc > 32) {
        // children with a non-negative score first, by score; the others (unexpanded, or a lost terminal's -0.05)
        // after them, by their original prior — the comparator below, as one key
        unsigned long long* key = reinterpret_cast<unsigned long long*>(C.expl);     // (filled with the exploration terms afterwards)
        __syncwarp();
        for (int i = V.lane; i < nc; i += 32) {
            const float sc = C.score[i];
            key[i] = sc < 0 ? (unsigned long long)gzp_fkey(C.prob_orig[i]) : ((1ULL << 32) | gzp_fkey(sc));
        }
        ordered = gzp_order_by_key(V, nc, key);
    }
#endif
    if (!ordered)
    gzp_order_children(V, nc, [&C](int a, int b) {
        const float sa = C.score[a], sb = C.score[b];     // -1 when unexpanded
        if (sa < 0 && sb < 0) return C.prob_orig[a] > C.prob_orig[b];
        return sa > sb;
    });
    // the exploration term of every child (an FP64 divide each) does not depend on the scan below: one child
    // per lane instead of all of them, one after the other, in every lane
    GZP_SYNC();
    for (int i = V.lane; i < nc; i += GZP_LANES) {
        const int traversals_i = (int)C.trav[i] + 1;
        const double inflight_i = C.to[i] >= 0 ? (double)C.inflight[i] : 0.0;
        C.expl[i] = (double)(puct_c * C.prob[i]) * sqrt_node_visits / (traversals_i + inflight_i);
    }
    GZP_SYNC();
#ifdef __CUDA_ARCH__
    // Lane-parallel form of the scan below for the common case: at most 32 children, no evaluation in
    // flight below this node and the root latch limiter inactive (so the scan draws no random numbers).
    // Lane r takes the child of rank r. The sequential scan keeps its running best as a FLOAT and compares
    // the next DOUBLE score against it; because rounding is monotonic, "child r replaces the best" is the
    // same as "score_r > max(initial, float(score_j) for eligible j < r)" — an exclusive prefix maximum —
    // and the final choice is the last child that replaced it.
    const bool latch_on = node_visits > 1000 && node_visits < 40000000 && depth == 0;
    if (nc <= 32 && !latch_on && V.E->EB == 1) {
        const unsigned full = 0xffffffffu;
        const int r = V.lane;
        const bool act = r < nc;
        const int pick = act ? order[r] : 0;
        const int to = act ? C.to[pick] : -1;
        const uint32_t ctrav = act ? C.trav[pick] : 0u;
        double child_score = prior_score;
        double exploration = act ? C.expl[pick] : 0.0;
        bool badfb = false, winret = false;
        int infl = 0;
        if (act && to >= 0) {
            infl = C.inflight[pick];
            child_score = C.score[pick];
            if ((C.flags[pick] & NF_FINALISED) != 0) {
                if (child_score > 0.99) {
                    if (depth > 0) winret = true;
                    else child_score *= 1.0f + puct_c;
                } else if (child_score < 0.01) {
                    badfb = true;
                } else {
                    exploration = 0.0;
                }
            }
        }
        if (!__any_sync(full, act && ctrav > 0 && infl > 0)) {
            const unsigned wm = __ballot_sync(full, act && winret);
            if (wm) {                                   // a proven win below the root: taken at once, lowest rank first
                const int p0 = order[__ffs(wm) - 1];
                pn[c->path_len] = n; pc[c->path_len] = p0; c->path_len++;
                if (info) info->to = C.to[p0];
                return p0;
            }
            const unsigned bm = __ballot_sync(full, act && badfb);
            const bool elig = act && !badfb;
            const double score = child_score + exploration;
            const float neg_inf = -3.402823466e+38f;
            float run = elig ? (float)score : neg_inf;   // inclusive prefix maximum of the float-rounded scores
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const float o = __shfl_up_sync(full, run, d);
                if (r >= d) run = fmaxf(run, o);
            }
            float before = __shfl_up_sync(full, run, 1);
            if (r == 0) before = neg_inf;
            before = fmaxf(before, -1.0f);               // float best_score = -1
            const unsigned win = __ballot_sync(full, elig && score > (double)before);
            int best_child = -1;
            if (win) best_child = order[31 - __clz(win)];
            else if (bm) best_child = order[31 - __clz(bm)];   // bad_fallback: the last proven loss
            if (best_child >= 0) {
                pn[c->path_len] = n; pc[c->path_len] = best_child; c->path_len++;
                if (info) info->to = C.to[best_child];
            }
            return best_child;
        }
    }
#endif
    for (int rank = 0; rank < nc; rank++) {
        const int pick = order[rank];
        if (C.flags[pick] & (CF_UNSELECTABLE | CF_BLOCKED)) { unselectables++; continue; }   // evaluator.cpp:384-394
        const int to = C.to[pick];
        const uint32_t ctrav = C.trav[pick];
        double child_score = prior_score;
        const double inflight = to >= 0 ? (double)C.inflight[pick] : 0.0;
        double exploration = C.expl[pick];
        if (to >= 0) {
            child_score = C.score[pick];
            const bool fin = (C.flags[pick] & NF_FINALISED) != 0;
            if (fin) {
                if (child_score > 0.99) {
                    if (depth > 0) {
                        pn[c->path_len] = n; pc[c->path_len] = pick; c->path_len++;
                        if (info) info->to = to;
                        return pick;
                    }
                    child_score *= 1.0f + puct_c;
                } else if (child_score < 0.01) {
                    bad_fallback = pick;
                    continue;
                } else {
                    exploration = 0.0;
                }
            }
            if ((fin || C.visits[pick] > 42) && child_score > best_child_score_actual) {
                best_child_score_actual = (float)child_score;
                best_child_score = pick;
            }
        }
        if (ctrav > 0 && inflight > 0) {
            const double discounted = inflight * (gzp_rng_get(c->rng_eval) + 0.5);
            child_score = (child_score * ctrav) / (ctrav + discounted);
        }
        const float limit_latch_root = 0.66f;
        const double score = child_score + exploration;
        if (node_visits > 1000 && node_visits < 40000000 && depth == 0 && gzp_rng_get(c->rng_eval) > 0.1) {
            if (ctrav > 16 && ctrav > node_visits * limit_latch_root) {
                if (best_fallback < 0 || score > best_fallback_score) {
                    best_fallback = pick;
                    best_fallback_score = (float)score;
                }
                continue;
            }
        }
        if (score > best_score) {
            best_child = pick;
            best_score = (float)score;
        }
    }
    if (best_child < 0) {
        if (best_fallback >= 0) best_child = best_child_score >= 0 ? best_child_score : best_fallback;
        else if (bad_fallback >= 0) {
            if (unselectables > 0 && info) info->yield_first = 1;
            best_child = bad_fallback;
        } else if (info) info->blocked = 1;
    }
    if (best_child >= 0) {
        pn[c->path_len] = n; pc[c->path_len] = best_child; c->path_len++;
        if (info) info->to = C.to[best_child];
    }
    return best_child;
}

// `nd` is node n's record as the caller read it (one 64-byte line); `preloaded_kid` (device warp path) this lane's
// child of n when the caller already holds it.
GZP_HDN int gzp_select_child(GameView& V, int n, const GzpNode& nd, GzpSelInfo* info = nullptr,
                             const GzpKid* preloaded_kid = nullptr) {
    const PuctEngine* E = V.E;
    GameCtl* c = V.c;
    const PuctConf& cf = gzp_conf(V);
    const int depth = c->path_len;
    if (c->path_len >= E->MAXD) { c->error = GZP_ERR_PATH; return -1; }
    const int nc = nd.n_nchild, c0 = nd.n_child0;
    const int lead = nd.n_lead;
    int32_t* pn = V.pn;
    int32_t* pc = V.pc;

    // setPuctConstant (evaluator.cpp:1300-1307)
    const uint32_t node_visits = nd.n_visits;
    const int node_inflight = nd.n_inflight;
    if (info) { info->to = -1; info->inflight = node_inflight; info->yield_first = 0; info->blocked = 0; }
    c->total_sel_nodes++;
    c->total_sel_children += (uint64_t)nc;
    V.work++;                 // the step's work budget counts every tree level, forced moves included
    if (nc == 1) {
        pn[c->path_len] = n; pc[c->path_len] = 0; c->path_len++;
        if (info) info->to = CHILD(c_to, c0);
        return 0;
    }
    float node_finals[GZP_MAX_ROLES];
#pragma unroll
    for (int r = 0; r < GZP_MAX_ROLES; r++) node_finals[r] = nd.final_[r];
    float puct_c;
    {
        const float base = 19652.0f;
        const float k = depth == 0 ? cf.puct_constant_root : cf.puct_constant;
        float pcst = gz_logf(((float)(1u + node_visits) + base) / base);   // std::log(float) == glibc logf
        pcst += k;
        puct_c = pcst;        // (the reference also stores it in the node, for its tree dump only)
    }
#ifdef __CUDA_ARCH__
    // warp path (children in registers) unless this is the root select that has to draw Dirichlet noise first
    if (nc <= 32 && V.ord && E->EB == 1 &&
        !(depth == 0 && cf.dirichlet_noise_pct >= 0 && !(nd.n_flags & NF_NOISE_SET) &&
          !(nd.score[lead < 0 ? 0 : lead] > 0.95))) {
        const GzpKid K = preloaded_kid ? *preloaded_kid : gzp_warp_load_kid(V, c0, nc, lead);
        int to = -1;
        const int pick = gzp_warp_select(V, K, nc, node_visits, node_finals[lead < 0 ? 0 : lead], puct_c, depth, &to);
        if (pick != GZP_WARP_PUNT) {
            if (pick >= 0) {
                pn[c->path_len] = n; pc[c->path_len] = pick; c->path_len++;
                if (info) info->to = to;
            }
            return pick;
        }
    } else if (nc > 32 && V.ord && E->EB == 1 && E->tt_cap == 0 && V.C.cap >= nc &&
               !(depth == 0 && cf.dirichlet_noise_pct >= 0 && !(nd.n_flags & NF_NOISE_SET) &&
                 !(nd.score[lead < 0 ? 0 : lead] > 0.95))) {
        int to = -1;
        const int pick = gzp_wide_select(V, n, nc, node_visits, node_finals[lead < 0 ? 0 : lead], puct_c, depth, &to);
        if (pick != GZP_WARP_PUNT && pick >= 0) {
            pn[c->path_len] = n; pc[c->path_len] = pick; c->path_len++;
            if (info) info->to = to;
            return pick;
        }
    }
#endif
    return gzp_select_child_generic(V, n, nd, info, puct_c);
}

#ifdef __CUDA_ARCH__
// backup on the warp path: lane i owns path element i (paths of at most 32 elements, no MCTS-prover
// finalisation). The walk from the leaf to the root is sequential only in the scores it hands upwards, and
// those change only at finalised nodes (which pass their own scores on instead, evaluator.cpp:578-583): the
// scores reaching element i are those of the nearest finalised element below it, else the leaf's. So every
// lane can fetch its node record and chosen-child record, apply the reference's update to its own node and
// write it back at once — two memory round trips for the whole path instead of two per level — and the copy
// of the child node's statistics kept in the parent's record (GzpChild) comes from the lane below by shuffle.
__device__ void gzp_backup_warp(GameView& V, const float* new_scores) {
    const PuctEngine* E = V.E;
    GameCtl* c = V.c;
    const int R = E->R;
    const int L = c->path_len;
    const int lane = V.lane;
    __syncwarp();
    const bool act = lane < L;
    const int n = act ? V.pn[lane] : 0;
    const int choice = act ? V.pc[lane] : -1;
    GzpNode nd;
    nd.n_parent = -1; nd.n_child0 = 0; nd.n_nchild = 0; nd.n_visits = 0; nd.n_inflight = 0; nd.n_flags = 0; nd.n_lead = 0; nd.n_depth = 0;
#pragma unroll
    for (int r = 0; r < GZP_MAX_ROLES; r++) { nd.score[r] = 0.0f; nd.final_[r] = 0.0f; }
    GzpNode* np = V.E->nodes + V.nb + n;
    if (act) nd = *np;
    GzpChild* kp = V.E->children + V.cb + nd.n_child0 + (choice >= 0 ? choice : 0);
    uint32_t ctrav = 0;
    float cprob = 0.0f;
    if (act && choice >= 0) { ctrav = kp->c_trav; cprob = kp->c_prob; }
    const int lead = nd.n_lead;
    const bool finalised = act && (nd.n_flags & NF_FINALISED) != 0;
    // the scores that reach this element
    const unsigned fin_mask = __ballot_sync(GZP_FULL, finalised);
    const unsigned below_fin = lane >= 31 ? 0u : (fin_mask & ~((2u << lane) - 1u));
    const int src = below_fin ? __ffs(below_fin) - 1 : lane;
    float sc[GZP_MAX_ROLES];
#pragma unroll
    for (int r = 0; r < GZP_MAX_ROLES; r++) {
        sc[r] = nd.score[r];
        if (r < R) {
            const float from = __shfl_sync(GZP_FULL, nd.score[r], src);
            const float ns = below_fin ? from : new_scores[r];
            if (!finalised) {
                float visits = (float)nd.n_visits;
                if (visits > 100000) visits = 100000 + 0.1f * (visits - 100000);
                sc[r] = (visits * sc[r] + ns) / (visits + 1.0f);
            }
        }
    }
    const uint32_t visits_u = nd.n_visits + 1u;
    const int inflight = nd.n_inflight > 0 ? nd.n_inflight - 1 : nd.n_inflight;
    // the statistics of the element below (the node this one's chosen child leads to), after its update
    const bool have_below = lane + 1 < L;
    float below_sc = 0.0f;
#pragma unroll
    for (int r = 0; r < GZP_MAX_ROLES; r++) {
        if (r < R) {
            const float v = __shfl_down_sync(GZP_FULL, sc[r], 1);
            if (r == (lead < 0 ? 0 : lead)) below_sc = v;
        }
    }
    const uint32_t below_visits = __shfl_down_sync(GZP_FULL, visits_u, 1);
    const int below_flags = __shfl_down_sync(GZP_FULL, nd.n_flags, 1);
    const int below_inflight = __shfl_down_sync(GZP_FULL, inflight, 1);
    bool need_norm = false;
    if (act) {
        if (!finalised) {
#pragma unroll
            for (int r = 0; r < GZP_MAX_ROLES; r++) if (r < R) np->score[r] = sc[r];
        }
        np->n_visits = visits_u;
        if (nd.n_inflight > 0) np->n_inflight = inflight;
        if (choice >= 0) {
            if (have_below) {
                kp->c_score = below_sc;
                kp->c_visits = below_visits;
                kp->c_flags = (uint8_t)(below_flags & NF_FINALISED);
                kp->c_inflight = (uint8_t)(below_inflight > 255 ? 255 : below_inflight);
            }
            kp->c_trav = ctrav + 1;
            if (visits_u > 23) {
                const float cur_score = lead == 1 ? sc[1] : lead == 2 ? sc[2] : lead == 3 ? sc[3] : sc[0];
                float apply, mn;
                if (cur_score > 0.3 && cur_score < 0.7) { apply = 0.995f; mn = 0.02f; }
                else if (cur_score > 0.15 && cur_score < 0.85) { apply = 0.9975f; mn = 0.03f; }
                else { apply = 0.9975f; mn = 0.10f; }
                float p = cprob;
                if (p > mn) {
                    p *= apply;
                    p = mn > p ? mn : p;
                    kp->c_prob = p;
                }
            }
        }
        need_norm = (visits_u % 100 == 0);
    }
    // normaliseX (node.h:192-214), once per 100 visits of a node: the generic loops, on the records
    unsigned norm = __ballot_sync(GZP_FULL, need_norm);
    while (norm) {
        const int index = __ffs(norm) - 1;
        norm &= norm - 1;
        __syncwarp();
        const int c0 = __shfl_sync(GZP_FULL, nd.n_child0, index), nc = __shfl_sync(GZP_FULL, nd.n_nchild, index);
        GzpChild* kids = V.E->children + V.cb + c0;
        float total = 0;
        for (int i = 0; i < nc; i++) total += kids[i].c_prob;
        __syncwarp();
        if (total > 1.17549435e-38f) {
            for (int i = lane; i < nc; i += 32) kids[i].c_prob /= total;
        } else {
            for (int i = lane; i < nc; i += 32) kids[i].c_prob = (float)(1.0 / nc);
        }
    }
    __syncwarp();
}
#endif

// backup (evaluator.cpp:519-656). Walking the path from the leaf up, every node's new statistics are also
// copied into its parent's child record (the next node on the walk), see GzpChild.
GZP_RARE void gzp_backup_generic(GameView& V, float* new_scores);
GZP_HD void gzp_backup(GameView& V, float* new_scores) {
#ifdef __CUDA_ARCH__
    if (!gzp_conf(V).backup_finalised && V.c->path_len <= 32) { gzp_backup_warp(V, new_scores); return; }
#endif
    gzp_backup_generic(V, new_scores);
}
GZP_RARE void gzp_backup_generic(GameView& V, float* new_scores) {
    const PuctEngine* E = V.E;
    GameCtl* c = V.c;
    const PuctConf& cf = gzp_conf(V);
    const int R = E->R;
    const int32_t* pn = V.pn;
    const int32_t* pc = V.pc;
    bool bp_finalised_only_once = cf.backup_finalised != 0;
    // statistics of the node handled in the previous iteration (the child of this one on the path)
    bool have_below = false;
    float below_sc[GZP_MAX_ROLES] = {0.0f, 0.0f, 0.0f, 0.0f};
    uint32_t below_visits = 0;
    int below_flags = 0, below_inflight = 0;
    for (int index = c->path_len - 1; index >= 0; index--) {
        const int n = pn[index];
        const int choice = pc[index];
        GzpNode& nd = V.E->nodes[V.nb + n];
        const int lead = nd.n_lead;
        const int c0n = nd.n_child0;
        GzpChild* kids = V.E->children + V.cb + c0n;
        if (choice >= 0 && have_below) {
            // before anything reads the children of n: the chosen child's node just changed
            GzpChild& k = kids[choice];
            k.c_score = below_sc[lead < 0 ? 0 : lead];
            k.c_visits = below_visits;
            k.c_flags = (uint8_t)(below_flags & NF_FINALISED);
            k.c_inflight = (uint8_t)(below_inflight > 255 ? 255 : below_inflight);
        }
        if (bp_finalised_only_once && !(nd.n_flags & NF_FINALISED) && lead >= 0) {
            bp_finalised_only_once = false;
            // forceFinalise
            float best_s = -1;
            int best = -1, win = -1;
            bool more = false;
            const int nc = nd.n_nchild;
            const bool direct = gzp_direct_stats(V.E);    // shared nodes: the records' copies may be stale
            for (int i = 0; i < nc; i++) {
                const GzpChild& k = kids[i];
                const bool fin = k.c_to >= 0 && (direct ? (NODE(n_flags, k.c_to) & NF_FINALISED) != 0 : (k.c_flags & NF_FINALISED) != 0);
                if (fin) {
                    const float s = direct ? NSCORE(k.c_to, lead) : k.c_score;
                    if (s > 0.99) { win = i; break; }
                    if (s > best_s) { best_s = s; best = i; }
                } else {
                    more = true;
                }
            }
            const int fc = win >= 0 ? win : (more ? -1 : best);
            if (fc >= 0) {
                const int to = kids[fc].c_to;
                for (int r = 0; r < R; r++) nd.score[r] = NSCORE(to, r);
                nd.n_flags |= NF_FINALISED;
            }
        }
        // everything this node needs, loaded in one go (through global pointers the compiler has to assume
        // that every store may alias the next load, so interleaved read-modify-writes cost a memory round
        // trip each); the arithmetic below is the reference's, in its order
        const int flags = nd.n_flags;
        const bool finalised = (flags & NF_FINALISED) != 0;
        uint32_t visits_u = nd.n_visits;
        int inflight = nd.n_inflight;
        float sc[GZP_MAX_ROLES];
#pragma unroll
        for (int r = 0; r < GZP_MAX_ROLES; r++) sc[r] = r < R ? nd.score[r] : 0.0f;
        uint32_t ctrav = 0;
        float cprob = 0.0f;
        if (choice >= 0) { ctrav = kids[choice].c_trav; cprob = kids[choice].c_prob; }
        if (finalised) {
            for (int r = 0; r < R; r++) new_scores[r] = sc[r];
        } else {
            for (int r = 0; r < R; r++) {
                float visits = (float)visits_u;
                if (visits > 100000) visits = 100000 + 0.1f * (visits - 100000);
                const float score = (visits * sc[r] + new_scores[r]) / (visits + 1.0f);
                sc[r] = score;
                nd.score[r] = score;
            }
        }
        visits_u += 1;
        nd.n_visits = visits_u;
        if (inflight > 0) { inflight -= 1; nd.n_inflight = inflight; }
        if (choice >= 0) {
            kids[choice].c_trav = ctrav + 1;
            if (visits_u > 23) {
                const float cur_score = sc[lead];
                float apply, mn;
                if (cur_score > 0.3 && cur_score < 0.7) { apply = 0.995f; mn = 0.02f; }
                else if (cur_score > 0.15 && cur_score < 0.85) { apply = 0.9975f; mn = 0.03f; }
                else { apply = 0.9975f; mn = 0.10f; }
                float p = cprob;
                if (p > mn) {
                    p *= apply;
                    p = mn > p ? mn : p;
                    kids[choice].c_prob = p;
                }
            }
        }
        if (visits_u % 100 == 0) {
            // normaliseX (node.h:192-214)
            const int nc = nd.n_nchild;
            float total = 0;
            for (int i = 0; i < nc; i++) total += kids[i].c_prob;
            if (total > 1.17549435e-38f) {
                for (int i = 0; i < nc; i++) kids[i].c_prob /= total;
            } else {
                for (int i = 0; i < nc; i++) kids[i].c_prob = (float)(1.0 / nc);
            }
        }
        have_below = true;
#pragma unroll
        for (int r = 0; r < GZP_MAX_ROLES; r++) below_sc[r] = sc[r];
        below_visits = visits_u;
        below_flags = flags;
        below_inflight = inflight;
    }
}

// converged (evaluator.cpp:1342-1362); sortedChildren = by visits of the child node desc, ties by
// policy_prob desc (node.cpp:312-343), unexpanded children count as 0 visits.
GZP_RARE bool gzp_converged_generic(GameView& V, int count, int n, int nc);
// `nd` = the root's record. On the device warp path the root's children are loaded into registers here and
// handed back through kid / have_kid for the select that follows.
GZP_HDN bool gzp_converged(GameView& V, int count, const GzpNode& nd, GzpKid* kid = nullptr, bool* have_kid = nullptr) {
    const int n = V.c->root;
    const int nc = nd.n_nchild;
    if (nc < 2) return true;
#ifdef __CUDA_ARCH__
    if (nc <= 32 && V.ord) {
        const GzpKid K = gzp_warp_load_kid(V, nd.n_child0, nc, nd.n_lead);
        if (kid) { *kid = K; *have_kid = true; }
        const int r = gzp_warp_converged(V, K, nc, count);
        if (r != GZP_WARP_PUNT) return r != 0;
    }
#endif
    return gzp_converged_generic(V, count, n, nc);
}
GZP_RARE bool gzp_converged_generic(GameView& V, int count, int n, int nc) {
    gzp_cache_fill(V, n);
    const ChildCache& C = V.C;
    int* order = C.order;
    bool ordered = false;
#ifdef __CUDA_ARCH__
    if (nc > 32) {
        unsigned long long* key = reinterpret_cast<unsigned long long*>(C.expl);
        __syncwarp();
        for (int i = V.lane; i < nc; i += 32)
            key[i] = ((unsigned long long)(unsigned)(int)C.visits[i] << 32) | gzp_fkey(C.prob[i]);
        ordered = gzp_order_by_key(V, nc, key);
    }
#endif
    if (!ordered)
    gzp_order_children(V, nc, [&C](int a, int b) {
        const int va = (int)C.visits[a], vb = (int)C.visits[b];     // 0 when unexpanded
        if (va == vb) return C.prob[a] > C.prob[b];
        return va > vb;
    });
    const int a = order[0], b = order[1];
    if (C.to[a] >= 0 && C.to[b] >= 0) {
        if (C.score[a] > C.score[b] && C.visits[a] > C.visits[b] + (uint32_t)count) return true;
    }
    return false;
}

// getTemperature (evaluator.cpp:1309-1322)
GZP_HD float gzp_temperature(const PuctConf& cf, int depth) {
    if (depth >= cf.depth_temperature_stop) return -1;
    float mult = 1.0f + ((depth - cf.depth_temperature_start) * cf.depth_temperature_increment);
    mult = mult > 1.0f ? mult : 1.0f;
    const float t = cf.temperature * mult;
    return t < cf.depth_temperature_max ? t : cf.depth_temperature_max;
}

// getProbabilities (evaluator.cpp:1161-1192): fills next_prob of every child.
GZP_RARE void gzp_get_probabilities(GameView& V, int n, float temperature, bool use_policy) {
    const int nc = NODE(n_nchild, n), c0 = NODE(n_child0, n);
    const float node_visits = (float)(NODE(n_visits, n) + 0.001 * nc);
    float total = 0.0f;
    for (int i = 0; i < nc; i++) {
        const float child_visits = CHILD(c_to, c0 + i) >= 0 ? (float)CHILD(c_trav, c0 + i) + 0.001f : 0.001f;
        float np = use_policy ? CHILD(c_prob, c0 + i) + 0.001f : child_visits / node_visits;
        np = (float)pow((double)np, (double)temperature);
        CHILD(c_next, c0 + i) = np;
        total += np;
    }
    for (int i = 0; i < nc; i++) CHILD(c_next, c0 + i) /= total;
}

// chooseTemperature (evaluator.cpp:1473-1510); iteration order = sortedChildren(.., next_prob).
GZP_RARE int gzp_choose_temperature(GameView& V, int n) {
    const PuctConf& cf = gzp_conf(V);
    const float temperature = gzp_temperature(cf, NODE(n_depth, n));
    if (temperature < 0) { gzp_cache_fill(V, n); return gzp_choose_top_visits(V, n); }
    const int root = V.c->root;
    gzp_get_probabilities(V, root, temperature, cf.dirichlet_noise_pct < 0 && NODE(n_visits, n) < 3);
    const float expected = (float)(gzp_rng_get(V.c->rng_eval) * cf.random_scale);
    const int nc = NODE(n_nchild, root);
    gzp_cache_fill(V, root);
    const ChildCache& C = V.C;
    int* order = C.order;
    gzp_order_children(V, nc, [&C](int a, int b) {
        const int va = (int)C.visits[a], vb = (int)C.visits[b];
        if (va == vb) return C.next[a] > C.next[b];
        return va > vb;
    });
    float seen = 0;
    for (int k = 0; k < nc; k++) {
        seen += C.next[order[k]];
        if (seen > expected) return order[k];
    }
    return order[nc - 1];
}

GZP_RARE int gzp_choose(GameView& V, int n) {
    const PuctConf& cf = gzp_conf(V);
    if (cf.choose == 1) return gzp_choose_temperature(V, n);
    gzp_cache_fill(V, n);
    return gzp_choose_top_visits(V, n);
}

// resetRootNode (evaluator.cpp:1019-1030)
GZP_RARE void gzp_reset_root_node(GameView& V) {
    const int n = V.c->root;
    const int nc = NODE(n_nchild, n), c0 = NODE(n_child0, n);
    for (int i = 0; i < nc; i++) {
        CHILD(c_prob, c0 + i) = CHILD(c_prob_orig, c0 + i);
        if (CHILD(c_trav, c0 + i) > 1u) CHILD(c_trav, c0 + i) = 1u;
    }
    NODE(n_flags, n) &= ~NF_NOISE_SET;
}

// ------------------------------------------------------------------------------------------------
// unique states (uniquestates.h): masked 64-bit hash -> saturating counter
GZP_HDN uint64_t gzp_state_hash(const PuctEngine* E, const uint32_t* st) {
    uint64_t h = 0xcbf29ce484222325ULL;
    for (int w = 0; w < E->SW; w++) {
        const uint64_t x = st[w] & E->uniq_mask[w];
        h ^= x + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2);
    }
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ULL; h ^= h >> 32;
    return h ? h : 1;
}

GZP_HDN int gzp_uniq_find(const PuctEngine* E, uint64_t key, bool insert, int lane) {
    (void)lane;
    if (!E->uniq_cap) return -1;
    uint32_t pos = (uint32_t)(key & (uint64_t)(E->uniq_cap - 1));
    for (int probe = 0; probe < 64; probe++) {
        unsigned long long cur = E->uniq_keys[pos];
        if (cur == key) return (int)pos;
        if (cur == 0ULL) {
            if (!insert) return -1;
#ifdef __CUDA_ARCH__
            unsigned long long prev = 0ULL;
            if (lane == 0) prev = atomicCAS(&E->uniq_keys[pos], 0ULL, (unsigned long long)key);
            prev = __shfl_sync(0xffffffffu, prev, 0);
            if (prev == 0ULL || prev == key) return (int)pos;
#else
            E->uniq_keys[pos] = key;
            return (int)pos;
#endif
        }
        pos = (pos + 1) & (uint32_t)(E->uniq_cap - 1);
    }
    return -1;
}

GZP_HDN bool gzp_is_unique(const PuctEngine* E, const uint32_t* st, int depth, int lane) {
    const int pos = gzp_uniq_find(E, gzp_state_hash(E, st), false, lane);
    if (pos < 0) return true;
    const int max_num_dupes = 1000;
    int allowed = max_num_dupes - 5 * depth;
    if (allowed < 2) allowed = 2;
    return E->uniq_counts[pos] < allowed;
}

GZP_HDN void gzp_uniq_add(const PuctEngine* E, const uint32_t* st, int lane) {
    const int pos = gzp_uniq_find(E, gzp_state_hash(E, st), true, lane);
    if (pos < 0) return;
#ifdef __CUDA_ARCH__
    if (lane == 0 && E->uniq_counts[pos] < 1000) atomicAdd(&E->uniq_counts[pos], 1);
    __syncwarp();
#else
    if (E->uniq_counts[pos] < 1000) E->uniq_counts[pos]++;
#endif
}

// ------------------------------------------------------------------------------------------------
// Transposition table (evaluator.cpp:144-173): keyed by the state under the transformer's hash mask; a hit needs
// the same game depth (the reference refuses others, they could close a cycle). Entries are added when a node is
// created (first node with a state wins, as std::unordered_map::emplace) and the table is rebuilt from the kept
// nodes when the tree is re-rooted (the reference erases the released nodes' keys; the two differ only when the
// same masked state sits at two depths of one tree).
GZP_HD bool gzp_masked_equal(const PuctEngine* E, const uint32_t* a, const uint32_t* b) {
    for (int w = 0; w < E->SW; w++)
        if ((a[w] ^ b[w]) & E->uniq_mask[w]) return false;
    return true;
}
GZP_HDN void gzp_tt_clear(GameView& V) {
    const PuctEngine* E = V.E;
    if (!E->tt_cap) return;
    GZP_SYNC();
    for (int i = V.lane; i < E->tt_cap; i += GZP_LANES) {
        E->tt_keys[(int64_t)V.g * E->tt_cap + i] = 0ULL;
        E->tt_node[(int64_t)V.g * E->tt_cap + i] = 0x7fffffff;
    }
    GZP_SYNC();
}
// single-threaded insert (node creation); `atomic` for the parallel rebuild after re-rooting
GZP_HDN void gzp_tt_insert(const PuctEngine* E, int g, int64_t nb, int node, bool atomic) {
    if (!E->tt_cap) return;
    uint64_t key = gzp_state_hash(E, E->n_state + (nb + node) * E->SW);
    if (!key) key = 1;                              // 0 marks an empty slot
    unsigned long long* keys = E->tt_keys + (int64_t)g * E->tt_cap;
    int32_t* nodes = E->tt_node + (int64_t)g * E->tt_cap;
    uint32_t pos = (uint32_t)(key & (uint64_t)(E->tt_cap - 1));
    for (int probe = 0; probe < E->tt_cap; probe++) {
#ifdef __CUDA_ARCH__
        if (atomic) {
            const unsigned long long prev = atomicCAS(&keys[pos], 0ULL, (unsigned long long)key);
            if (prev == 0ULL || prev == key) { atomicMin(&nodes[pos], node); return; }
        } else
#endif
        {
            (void)atomic;
            if (keys[pos] == 0ULL) { keys[pos] = key; nodes[pos] = node; return; }
            if (keys[pos] == key) { if (node < nodes[pos]) nodes[pos] = node; return; }
        }
        pos = (pos + 1) & (uint32_t)(E->tt_cap - 1);
    }
}
// lookupNode: the node holding `state` (masked) at `depth`, or -1
GZP_HDN int gzp_tt_find(const GameView& V, const uint32_t* state, int depth) {
    const PuctEngine* E = V.E;
    if (!E->tt_cap) return -1;
    uint64_t key = gzp_state_hash(E, state);
    if (!key) key = 1;
    const unsigned long long* keys = E->tt_keys + (int64_t)V.g * E->tt_cap;
    const int32_t* nodes = E->tt_node + (int64_t)V.g * E->tt_cap;
    uint32_t pos = (uint32_t)(key & (uint64_t)(E->tt_cap - 1));
    for (int probe = 0; probe < E->tt_cap; probe++) {
        if (keys[pos] == 0ULL) return -1;
        if (keys[pos] == key) {
            const int n = nodes[pos];
            if (n < V.c->n_nodes && gzp_masked_equal(E, state, NSTATE(n)))
                return NODE(n_depth, n) == depth ? n : -1;
        }
        pos = (pos + 1) & (uint32_t)(E->tt_cap - 1);
    }
    return -1;
}

// ------------------------------------------------------------------------------------------------
// SelfPlayManager::createSample (selfplaymanager.cpp:72-119) into the game's pending list.
GZP_RARE void gzp_create_sample(GameView& V, int n) {
    const PuctEngine* E = V.E;
    GameCtl* c = V.c;
    if (c->n_samples >= E->MAXPEND) { c->error = GZP_ERR_SAMPLES; return; }
    int32_t* rec = E->pend + ((int64_t)V.g * E->MAXPEND + c->n_samples) * E->srec;
    c->n_samples++;
    const int nc = NODE(n_nchild, n), c0 = NODE(n_child0, n);
    const int lead = NODE(n_lead, n);
    rec[SR_READY] = 0;
    rec[SR_DEPTH] = NODE(n_depth, n);
    rec[SR_SLOT] = V.g;
    rec[SR_VISITS] = (int32_t)NODE(n_visits, n);
    rec[SR_LEAD] = lead;
    rec[SR_N_PREV] = c->root_has_prev ? 1 : 0;
    rec[SR_N_CHILDREN] = nc;
    float* fr = reinterpret_cast<float*>(rec);
    for (int r = 0; r < GZP_MAX_ROLES; r++) {
        fr[SR_PUCT_SCORE + r] = r < E->R ? NSCORE(n, r) : 0.0f;
        rec[SR_OTHER_MOVES + r] = r < E->R ? NOTHER(n, r) : 0;
    }
    int32_t* p = rec + SR_HEADER;
    const uint32_t* st = NSTATE(n);
    for (int w = V.lane; w < E->SW; w += GZP_LANES) p[w] = (int32_t)st[w];
    p += E->SW;
    const uint32_t* ps = E->root_prev_state + (int64_t)V.g * E->SW;
    for (int w = V.lane; w < E->SW; w += GZP_LANES) p[w] = c->root_has_prev ? (int32_t)ps[w] : 0;
    p += E->SW;
    for (int i = V.lane; i < nc; i += GZP_LANES) {
        p[2 * i] = CHILD(c_move, c0 + i);
        reinterpret_cast<float*>(p)[2 * i + 1] = CHILD(c_next, c0 + i);
    }
    GZP_SYNC();
}

// SelfPlay::addSamples (selfplay.cpp:244-288): stamp the game's samples and publish them.
GZP_RARE void gzp_add_samples(GameView& V, const float* final_scores, int game_depth) {
    const PuctEngine* E = V.E;
    GameCtl* c = V.c;
    bool fp0 = false, fp1 = false;
    for (int r = 0; r < E->R; r++) {
        const float fs = final_scores[r];
        if (c->has_resigned) {
            if (!fp0 && c->resign0_set) {
                const float s = c->resign0_scores[r];
                if ((s < E->conf.resign0_score_probability * 1.05) && fs > 0.49) { fp0 = true; c->false_positive_resigns0++; }
            }
            if (!fp1 && c->resign1_set) {
                const float s = c->resign1_scores[r];
                if ((s < E->conf.resign1_score_probability * 1.05) && fs > 0.49) { fp1 = true; c->false_positive_resigns1++; }
            }
        }
    }
    for (int k = 0; k < c->n_samples; k++) {
        int32_t* rec = E->pend + ((int64_t)V.g * E->MAXPEND + k) * E->srec;
        float* fr = reinterpret_cast<float*>(rec);
        for (int r = 0; r < GZP_MAX_ROLES; r++) fr[SR_FINAL_SCORE + r] = r < E->R ? final_scores[r] : 0.0f;
        rec[SR_GAME_LENGTH] = game_depth;
        rec[SR_MATCH] = c->match_count;
        rec[SR_HAS_RESIGNED] = c->has_resigned;
        rec[SR_RESIGN_FP] = (fp0 || fp1) ? 1 : 0;
        rec[SR_START_DEPTH] = c->starting_sample_depth;
        rec[SR_READY] = 1;
#ifdef __CUDA_ARCH__
        unsigned long long slot = 0ULL;
        __syncwarp();
        if (V.lane == 0) slot = atomicAdd(E->ready_count, 1ULL);
        slot = __shfl_sync(0xffffffffu, slot, 0);
#else
        const unsigned long long slot = (*E->ready_count)++;
#endif
        if (slot < (unsigned long long)E->READYCAP) {
            int32_t* dst = E->ready + (int64_t)slot * E->srec;
            for (int w = V.lane; w < E->srec; w += GZP_LANES) dst[w] = rec[w];
        } else {
            // the caller did not fetch for too long: the record is lost, and counted (gz_puct_totals: dropped_samples)
#ifdef __CUDA_ARCH__
            if (V.lane == 0) atomicAdd(E->ready_count + 1, 1ULL);
#else
            E->ready_count[1]++;
#endif
        }
        GZP_SYNC();
        c->total_samples++;
    }
    c->n_samples = 0;
}

// SelfPlay::resign (selfplay.cpp:45-74)
GZP_RARE bool gzp_resign(GameView& V, int n) {
    const PuctEngine* E = V.E;
    GameCtl* c = V.c;
    const float score = NSCORE(n, NODE(n_lead, n));
    if (c->can_resign0 && score < E->conf.resign0_score_probability) {
        c->has_resigned = 1;
        c->resign0_set = 1;
        for (int r = 0; r < E->R; r++) c->resign0_scores[r] = NSCORE(n, r);
    } else if (c->can_resign1 && score < E->conf.resign1_score_probability) {
        c->has_resigned = 1;
        c->resign1_set = 1;
        for (int r = 0; r < E->R; r++) c->resign1_scores[r] = NSCORE(n, r);
    }
    return c->has_resigned != 0;
}

// ------------------------------------------------------------------------------------------------
// Re-rooting (replaces fastApplyMove's ref-count release of the sibling subtrees,
// evaluator.cpp:888-943): the subtree under `new_root` is compacted IN PLACE to the front of the
// game's arenas. Nodes are stored in creation order, so parent index < child index, new index <= old
// index, and child blocks keep their relative order: both arrays can be compacted left to right.
// The work is written for a cooperating group of `T` threads (`tid` in [0,T), `sync()` a group
// barrier, `scan()` an exclusive group prefix sum that also returns the total): a whole CTA on the
// device (puct_reroot_kernel), a single thread on the host build.
struct GzpSeq {   // trivial group of one thread
    int total_a, total_b;
    GZP_HD void sync() {}
    GZP_HD void scan2(int a, int b, int& ea, int& eb, int& ta, int& tb) { ea = 0; eb = 0; ta = a; tb = b; }
};

template <typename Group>
GZP_HDN void gzp_reroot_compact(GameView& V, int new_root, int tid, int T, Group& grp) {
    const PuctEngine* E = V.E;
    GameCtl* c = V.c;
    const int64_t nb = V.nb, cb = V.cb;
    const int N = c->n_nodes, NC = c->n_children;
    const int r = new_root;
    // keep the state of the old root: it is the previous state of the new root (samples, encoder)
    {
        uint32_t* ps = E->root_prev_state + (int64_t)V.g * E->SW;
        const uint32_t* st = E->n_state + (nb + c->root) * E->SW;
        for (int w = tid; w < E->SW; w += T) ps[w] = st[w];
    }
    // ---- pass 0 (transposition tables only): the tree is a DAG, a shared node is kept when ANY kept node links to
    // it (the reference's ref counts, evaluator.cpp:113-140) and a linking parent may have been created after the
    // node it links to. Marks spread along the child links, one depth level per round at worst.
    const bool dag = E->tt_cap != 0;
    if (dag) {
        for (int i = tid; i < N; i += T) E->n_newidx[nb + i] = i == r ? 0 : -1;
        grp.sync();
        for (;;) {
            int changed = 0;
            for (int k = tid; k < NC; k += T) {
                const int to = E->children[cb + k].c_to;
                if (to >= 0 && E->n_newidx[nb + to] < 0 && E->n_newidx[nb + E->c_owner[cb + k]] >= 0) {
                    E->n_newidx[nb + to] = 0;
                    changed = 1;
                }
            }
            int e0, e1, total, t1;
            grp.scan2(changed, 0, e0, e1, total, t1);
            if (!total) break;
        }
    }
    // ---- pass 1: reachability + new indices, chunk by chunk in index order
    int node_base = 0, child_base = 0;
    for (int b = 0; b < N; b += T) {
        const int i = b + tid;
        int reach = 0, nch = 0;
        if (i < N) {
            if (dag) {
                reach = E->n_newidx[nb + i] >= 0;
            } else {
                int q = i;
                while (q > r && q >= b) q = E->nodes[nb + q].n_parent;
                if (q == r) reach = 1;
                else if (q > r) reach = E->n_newidx[nb + q] >= 0;     // q < b: decided by an earlier chunk
            }
            if (reach) nch = E->nodes[nb + i].n_nchild;
        }
        int en, ec, tn, tc;
        grp.scan2(reach, nch, en, ec, tn, tc);
        if (i < N) {
            E->n_newidx[nb + i] = reach ? node_base + en : -1;
            E->n_newc0[nb + i] = reach ? child_base + ec : -1;
        }
        node_base += tn;
        child_base += tc;
        grp.sync();
    }
    // ---- pass 2: child records, in place, left to right (new position <= old position)
    for (int b = 0; b < NC; b += T) {
        const int k = b + tid;
        int dst = -1, own = 0;
        GzpChild rec;
        if (k < NC) {
            const int o = E->c_owner[cb + k];
            const int no = E->n_newidx[nb + o];
            if (no >= 0) {
                dst = E->n_newc0[nb + o] + (k - E->nodes[nb + o].n_child0);
                rec = E->children[cb + k];
                if (rec.c_to >= 0) rec.c_to = E->n_newidx[nb + rec.c_to];
                own = no;
            }
        }
        grp.sync();
        if (dst >= 0) {
            E->children[cb + dst] = rec;
            E->c_owner[cb + dst] = own;
        }
        grp.sync();
    }
    // ---- pass 3: node records, in place
    for (int b = 0; b < N; b += T) {
        const int i = b + tid;
        int dst = -1;
        GzpNode rec;
        int ot[GZP_MAX_ROLES];
        if (i < N) {
            dst = E->n_newidx[nb + i];
            if (dst >= 0) {
                rec = E->nodes[nb + i];
                const int p0 = rec.n_parent;
                rec.n_parent = (i == r || p0 < 0) ? -1 : E->n_newidx[nb + p0];    // a shared node's creator may be gone: -1
                rec.n_child0 = rec.n_nchild == 0 ? -1 : E->n_newc0[nb + i];
                for (int k = 0; k < GZP_MAX_ROLES; k++) ot[k] = E->n_other[(nb + i) * GZP_MAX_ROLES + k];
            }
        }
        grp.sync();
        if (dst >= 0) {
            E->nodes[nb + dst] = rec;
            for (int k = 0; k < GZP_MAX_ROLES; k++) E->n_other[(nb + dst) * GZP_MAX_ROLES + k] = ot[k];
        }
        grp.sync();
    }
    // ---- pass 4: states (rows of SW words), in place; every thread moves whole rows in index order
    for (int b = 0; b < N; b += T) {
        const int i = b + tid;
        const int dst = i < N ? E->n_newidx[nb + i] : -1;
        // rows are small (SW words); a thread copies its row into registers chunk-wise
        for (int w0 = 0; w0 < E->SW; w0 += 8) {
            uint32_t tmp[8];
            const int nw = E->SW - w0 < 8 ? E->SW - w0 : 8;
            if (dst >= 0) for (int w = 0; w < nw; w++) tmp[w] = E->n_state[(nb + i) * E->SW + w0 + w];
            grp.sync();
            if (dst >= 0) for (int w = 0; w < nw; w++) E->n_state[(nb + dst) * E->SW + w0 + w] = tmp[w];
            grp.sync();
        }
    }
    if (dag) {
        // the table is rebuilt from the kept nodes (the reference erases the released nodes' keys one by one)
        for (int i = tid; i < E->tt_cap; i += T) {
            E->tt_keys[(int64_t)V.g * E->tt_cap + i] = 0ULL;
            E->tt_node[(int64_t)V.g * E->tt_cap + i] = 0x7fffffff;
        }
        grp.sync();
        for (int i = tid; i < node_base; i += T) gzp_tt_insert(E, V.g, nb, i, true);
        grp.sync();
    }
    if (tid == 0) {
        // a tree's new root is its first kept node; in a DAG shared nodes created before it may precede it
        c->root = E->n_newidx[nb + r];
        c->n_nodes = node_base;
        c->n_children = child_base;
        c->game_depth++;
        c->root_has_prev = 1;
    }
    grp.sync();
}

GZP_HDN void gzp_reroot_host(GameView& V, int new_root) {
    GzpSeq grp;
    gzp_reroot_compact(V, new_root, 0, 1, grp);
}

GZP_HDN void gzp_request_reroot(GameView& V, int new_root) {
    GameCtl* c = V.c;
    c->req = REQ_REROOT;
    c->req_parent = new_root;
    c->req_child = -1;
    c->phase = PH_AFTER_MOVE;
}

// ------------------------------------------------------------------------------------------------
// Controller: the reference's coroutine per game (SelfPlay::playGamesForever -> playOnce ->
// collectSamples / runToEnd -> PuctEvaluator::onNextMove -> playoutMain -> treePlayout), turned
// into a resumable state machine. gzp_step_pre runs until the game needs the batch pipeline
// (a propnet transition and possibly a network evaluation); gzp_step_post consumes the propnet
// outputs and creates the node.
enum { CMD_NONE = 0, CMD_RESET = 1, CMD_APPLY = 2, CMD_SEARCH = 3, CMD_BALANCE = 4 };
// wall-clock limits of playoutMain (evaluator.cpp:794,820-826: end_time, think_time) live on the host, which
// tells a running search to stop: GZP_STOP_NOW ("hard" limit, or think time x multiplier without convergence),
// GZP_STOP_IF_CONVERGED (think time elapsed). Checked where the reference checks its clock: between playouts.
enum { GZP_STOP_NONE = 0, GZP_STOP_NOW = 1, GZP_STOP_IF_CONVERGED = 2 };

GZP_HDN void gzp_request_expand(GameView& V, int parent, int child, int purpose) {
    const PuctEngine* E = V.E;
    GameCtl* c = V.c;
    c->req = REQ_EXPAND;
    c->req_parent = parent;
    c->req_child = child;
    c->req_purpose = purpose;
    uint32_t* xin = E->x_in_state + (int64_t)V.g * E->SW;
    const uint32_t* st = NSTATE(parent);
    for (int w = V.lane; w < E->SW; w += GZP_LANES) xin[w] = st[w];      // (hexLG13: 932 words per state)
    GZP_SYNC();
    const int lead = NODE(n_lead, parent);
    const int ci = NODE(n_child0, parent) + child;
    for (int r = 0; r < E->R; r++)
        E->x_moves[(int64_t)V.g * E->R + r] = (int16_t)(r == lead ? CHILD(c_move, ci) : NOTHER(parent, r));
}

GZP_HDN void gzp_request_root(GameView& V, const uint32_t* state) {
    const PuctEngine* E = V.E;
    GameCtl* c = V.c;
    c->req = REQ_ROOT;
    c->req_parent = -1;
    c->req_child = -1;
    c->req_purpose = PUR_ROOT;
    uint32_t* xin = E->x_in_state + (int64_t)V.g * E->SW;
    for (int w = V.lane; w < E->SW; w += GZP_LANES) xin[w] = state[w];
    GZP_SYNC();
    for (int r = 0; r < E->R; r++) E->x_moves[(int64_t)V.g * E->R + r] = -1;   // identity transition
}

GZP_HDN void gzp_finish_playout(GameView& V, const GzpNode& leaf) {
    const PuctEngine* E = V.E;
    GameCtl* c = V.c;
    float scores[GZP_MAX_ROLES];
    if (leaf.n_flags & NF_FINALISED) c->playouts_finals++;
    for (int r = 0; r < E->R; r++) scores[r] = leaf.score[r];
    gzp_backup(V, scores);
    c->num_tree_playouts++;
    c->total_playouts++;
    c->in_playout = 0;
    V.work += 2;              // the fixed cost of a playout (convergence check, backup) in tree levels, roughly
}

GZP_HDN void gzp_path_push(GameView& V, int n, int choice) {
    const PuctEngine* E = V.E;
    GameCtl* c = V.c;
    if (c->path_len >= E->MAXD) { c->error = GZP_ERR_PATH; return; }
    V.pn[c->path_len] = n;
    V.pc[c->path_len] = choice;
    c->path_len++;
}

// treePlayout (evaluator.cpp:658-720), resumable. Returns true when blocked on an expansion.
// `root_nd` / `root_kid`: the root's record (and, on the device warp path, its children) when the caller has
// just read them and this call starts a playout at the root.
GZP_HDN bool gzp_playout_continue(GameView& V, const GzpNode* root_nd = nullptr, const GzpKid* root_kid = nullptr) {
    GameCtl* c = V.c;
    while (true) {
        const int cur = c->cur;
        const GzpNode nd = root_nd ? *root_nd : V.E->nodes[V.nb + cur];       // one 64-byte line per tree level
        if ((nd.n_flags & (NF_FORCE_TERMINAL | NF_FINALISED)) || nd.n_nchild == 0) {
            gzp_path_push(V, cur, -1);
            gzp_finish_playout(V, nd);
            return false;
        }
        GzpSelInfo si;
        const int child = gzp_select_child(V, cur, nd, &si, root_nd ? root_kid : nullptr);
        root_nd = nullptr;
        if (child < 0 || c->error) { c->error = c->error ? c->error : GZP_ERR_PATH; return true; }
        if (si.to < 0) {
            gzp_request_expand(V, cur, child, PUR_PLAYOUT);
            return true;
        }
        NODE(n_inflight, cur) = si.inflight + 1;
        c->cur = si.to;
    }
}

GZP_HDN void gzp_start_search(GameView& V, int evals, int conf_sel) {
    GameCtl* c = V.c;
    c->conf_sel = conf_sel;
    c->max_evaluations = evals;
    c->num_evaluations = 0;
    c->num_tree_playouts = 0;
    c->playouts_finals = 0;
    c->in_playout = 0;
    c->stop_request = GZP_STOP_NONE;
    c->phase = PH_SEARCH;
}

GZP_RARE void gzp_begin_game(GameView& V) {
    const PuctEngine* E = V.E;
    GameCtl* c = V.c;
    c->match_count++;
    c->n_samples = 0;
    c->has_resigned = 0;
    const double r = gzp_rng_get(c->rng_sp);
    c->can_resign0 = r > E->conf.resign0_pct;
    c->can_resign1 = r > E->conf.resign1_pct;
    c->resign0_set = c->resign1_set = 0;
    // PuctEvaluator::reset(0)
    c->n_nodes = 0;
    c->n_children = 0;
    c->game_depth = 0;
    c->root = -1;
    c->root_has_prev = 0;
    c->in_run_to_end = 0;
    c->in_playout = 0;
    c->conf_sel = 0;
    gzp_tt_clear(V);
    if (c->start_pending) {
        // staggered start (measurement aid): this slot's game begins from a given position, at depth 0 like
        // the reference's establishRoot(initial_state) would with that state handed in as the initial one
        c->start_pending = 0;
        gzp_request_root(V, E->start_state + (int64_t)V.g * E->SW);
    } else {
        gzp_request_root(V, E->init_state);
    }
}

// A self-play game that ran into one of the engine's per-game capacities (tree arena, path length, pending
// samples — the reference has no such limits) is abandoned: its samples are discarded, the restart is counted
// and the slot begins a new game. Rule violations (GZP_ERR_SIMULTANEOUS, GZP_ERR_ROOT_TERMINAL) and player-mode
// errors still halt the slot.
GZP_HD bool gzp_restart_after_error(GameView& V) {
    GameCtl* c = V.c;
    const int e = c->error;
    const bool recoverable = e == GZP_ERR_NODES || e == GZP_ERR_CHILDREN || e == GZP_ERR_PATH || e == GZP_ERR_SAMPLES;
    if (c->mode != 0 || !recoverable) return false;
    c->last_error = e;
    c->error = GZP_ERR_NONE;
    c->error_restarts++;
    c->n_samples = 0;
    c->req = REQ_NONE;
    c->eval_node = -1;
    c->resume = 0;
    c->in_playout = 0;
    c->path_len = 0;
    V.E->x_need_eval[(int64_t)V.g * V.E->EB] = 0;
    c->phase = PH_NEW_GAME;
    return true;
}


// playoutMain's list of break conditions (evaluator.cpp:782-841); true: the search is over.
GZP_HD bool gzp_search_over(GameView& V, const GzpNode& root_nd, GzpKid* root_kid, bool* have_kid) {
    const PuctEngine* E = V.E;
    GameCtl* c = V.c;
    const PuctConf& cf = gzp_conf(V);
    const bool is_converged = gzp_converged(V, cf.converged_visits, root_nd, root_kid, have_kid);
    const int max_non_converged = (int)(c->max_evaluations * cf.evaluation_multiplier_to_convergence);
    const int max_tree_playouts = max_non_converged > (1 << 28) ? (1 << 30) : 4 * max_non_converged;
    if (c->stop_request == GZP_STOP_NOW || (c->stop_request == GZP_STOP_IF_CONVERGED && is_converged)) return true;
    if ((root_nd.n_flags & NF_FINALISED) && c->num_tree_playouts > 100) return true;
    if (is_converged && c->num_tree_playouts > max_tree_playouts) return true;
    // arena nearly full ("max nodes"); every coroutine of a search with workers may still add a node before the next check
    if (c->n_nodes + 4 * E->EB > E->NCAP || c->n_children + E->max_actions * E->EB + 4 > E->CCAP) return true;
    if (is_converged && c->num_evaluations > c->max_evaluations) return true;
    if (!is_converged && c->num_evaluations > max_non_converged) return true;
    // Deviation (documented in DESIGN.md): the reference loops forever when a search only reaches terminal /
    // finalised nodes and the root never 'converges' (no evaluation is ever counted); a game slot must not be
    // lost, so such a search stops here.
    if ((long long)c->num_tree_playouts > 64LL * max_tree_playouts + 10000LL) return true;
    return false;
}

// ------------------------------------------------------------------------------------------------
// Intra-tree workers (PuctConfig.batch_size > 1, a Player's search). The reference runs playoutMain and
// batch_size - 1 playoutWorker coroutines (greenlets) over ONE tree; a coroutine runs until it blocks in
// NetworkScheduler::evaluate (its leaf joins the batch) or yields, and NetworkScheduler::mainLoop
// (scheduler.cpp:131-203) picks the next one from a FIFO of runnables, hands the batch to the network when it is
// full or nobody can run, answers the requestors in order and re-queues them, then the yielders. The order in
// which coroutines touch the tree decides every visit count, so that schedule is restated here exactly: each
// coroutine is a resumable state machine (GzpCo), the queues live in GzpSched. One difference in mechanics, none
// in order: a coroutine that expands a child waits for the engine's propnet pass while keeping the controller
// (nobody else runs in between — in the reference the state machine call is synchronous).
//
// All lanes of the warp run this code redundantly on one copy of the queues in global memory: every read of
// queue state is bracketed by warp barriers before anything is written back.
enum { CO_R_PROPNET = 1, CO_R_BLOCKED, CO_R_YIELD, CO_R_DONE };

GZP_HD int gzp_sync_load(const int32_t* p) {
    GZP_SYNC();
    const int v = *p;
    GZP_SYNC();
    return v;
}
GZP_HD void gzp_queue_push(int32_t* q, int32_t* n, int k) {
    const int m = gzp_sync_load(n);
    q[m] = k;
    *n = m + 1;
    GZP_SYNC();
}

GZP_HD void gzp_co_bind(GameView& V, int k) {     // the controller's playout context becomes coroutine k's
    const PuctEngine* E = V.E;
    GzpSched& S = E->sched[V.g];
    V.pn = E->co_path_node + ((int64_t)V.g * (E->EB + 1) + k) * E->MAXD;
    V.pc = E->co_path_choice + ((int64_t)V.g * (E->EB + 1) + k) * E->MAXD;
    V.c->cur = S.co[k].cur;
    V.c->path_len = S.co[k].path_len;
}
GZP_HD void gzp_co_save(GameView& V, int k) {
    GzpSched& S = V.E->sched[V.g];
    GZP_SYNC();
    S.co[k].cur = V.c->cur;
    S.co[k].path_len = V.c->path_len;
    GZP_SYNC();
}

GZP_RARE void gzp_sched_start(GameView& V) {       // Player::puctPlayerMove: one runnable, the search itself
    GzpSched& S = V.E->sched[V.g];
    GZP_SYNC();
    S.n_run = 1; S.run[0] = 0; S.n_req = 0; S.n_yld = 0; S.current = -1;
    S.worker_count = 0; S.do_playouts = 1; S.flush_wait = 0; S.active = 1;
    for (int k = 0; k <= GZP_MAX_EB; k++) { S.co[k].st = CO_DEAD; S.co[k].req_parent = -1; S.co[k].req_child = -1; S.co[k].kind = CO_KIND_LOOP; }
    S.co[0].st = S.balance ? CO_BAL_START : CO_MAIN_START;
    GZP_SYNC();
}

// the end of a playout for coroutine k: back to the top of its loop
GZP_HD int gzp_co_loop_top(int k) { return k == 0 ? CO_MAIN_TOP : CO_WORKER_TOP; }

// Runs coroutine k until it gives up the controller. Returns CO_R_*.
GZP_RARE int gzp_co_run(GameView& V, int k) {
    const PuctEngine* E = V.E;
    GameCtl* c = V.c;
    GzpSched& S = E->sched[V.g];
    GzpCo& co = S.co[k];
    gzp_co_bind(V, k);
    for (int guard = 0; guard < (1 << 20); guard++) {
        if (c->error) return CO_R_DONE;
        const int st = gzp_sync_load(&co.st);
        switch (st) {
        case CO_MAIN_START: {
            // onNextMove (evaluator.cpp:1032-1069): the workers are queued behind whatever is runnable
            const int root = c->root;
            const int B = gzp_conf(V).batch_size < E->EB ? gzp_conf(V).batch_size : E->EB;
            if (B > 1 && !gzp_is_finalised(V, root) && (c->max_evaluations < 0 || c->max_evaluations > 100)) {
                for (int w = 1; w < B; w++) {
                    S.co[w].st = CO_WORKER_TOP;
                    gzp_queue_push(S.run, &S.n_run, w);
                }
                S.worker_count = B - 1;
            }
            co.st = c->max_evaluations != 0 ? CO_MAIN_TOP : CO_MAIN_COLLECT;
            break;
        }
        case CO_MAIN_TOP: {
            const GzpNode root_nd = V.E->nodes[V.nb + c->root];
            bool have_kid = false;
            GzpKid kid;
            if (gzp_search_over(V, root_nd, &kid, &have_kid)) { co.st = CO_MAIN_COLLECT; break; }
            c->cur = c->root;
            c->path_len = 0;
            co.st = CO_DESCEND;
            break;
        }
        case CO_MAIN_COLLECT: {
            // do_playouts = false; while (worker_count > 0) yield();
            S.do_playouts = 0;
            if (gzp_sync_load(&S.worker_count) > 0) { gzp_co_save(V, k); return CO_R_YIELD; }
            co.st = CO_DEAD;
            return CO_R_DONE;
        }
        case CO_WORKER_TOP: {
            if (!gzp_sync_load(&S.do_playouts)) { co.st = CO_DEAD; S.worker_count = gzp_sync_load(&S.worker_count) - 1; return CO_R_DONE; }
            if (c->num_tree_playouts % 10000 == 0) { co.st = CO_WORKER_RESUMED; gzp_co_save(V, k); return CO_R_YIELD; }
            co.st = CO_WORKER_RESUMED;
            break;
        }
        case CO_WORKER_RESUMED: {
            if (gzp_is_finalised(V, c->root)) { co.st = CO_DEAD; S.worker_count = gzp_sync_load(&S.worker_count) - 1; return CO_R_DONE; }
            c->cur = c->root;
            c->path_len = 0;
            co.st = CO_DESCEND;
            break;
        }
        case CO_DESCEND: {
            // treePlayout's loop (evaluator.cpp:666-707)
            const int cur = c->cur;
            const GzpNode nd = V.E->nodes[V.nb + cur];
            if ((nd.n_flags & (NF_FORCE_TERMINAL | NF_FINALISED)) || nd.n_nchild == 0) {
                gzp_path_push(V, cur, -1);
                gzp_finish_playout(V, nd);
                if (co.kind == CO_KIND_ONCE) { co.st = CO_DEAD; S.worker_count = gzp_sync_load(&S.worker_count) - 1; return CO_R_DONE; }
                co.st = gzp_co_loop_top(k);
                break;
            }
            GzpSelInfo si;
            const int child = gzp_select_child(V, cur, nd, &si);
            if (c->error) return CO_R_DONE;
            if (child < 0) {                    // nothing selectable right now: yield, then select again
                gzp_co_save(V, k);
                return CO_R_YIELD;
            }
            if (si.yield_first) {               // the yield inside selectChild; the chosen child is an expanded node
                co.sel_to = si.to;
                co.st = CO_SELECT_RESUMED;
                gzp_co_save(V, k);
                return CO_R_YIELD;
            }
            if (si.to < 0) {
                // expandChild: the child is unselectable from here until this coroutine has its node
                co.req_parent = cur;
                co.req_child = child;
                co.st = CO_AFTER_PROPNET;
                gzp_request_expand(V, cur, child, PUR_PLAYOUT);
                gzp_co_save(V, k);
                return CO_R_PROPNET;
            }
            NODE(n_inflight, cur) = si.inflight + 1;
            c->cur = si.to;
            break;
        }
        case CO_SELECT_RESUMED: {
            const int cur = c->cur;
            NODE(n_inflight, cur) = NODE(n_inflight, cur) + 1;
            c->cur = co.sel_to;
            co.st = CO_DESCEND;
            break;
        }
        case CO_AFTER_PROPNET: {
            // gzp_step_post has looked the state up / created the node
            const int N = c->new_node;
            co.node = N;
            co.st = CO_AFTER_EVAL;
            if (c->eval_node >= 0) {            // createNode -> NetworkScheduler::evaluate: join the batch and wait
                c->eval_node = -1;
                co.evaluated = 1;
                gzp_queue_push(S.req, &S.n_req, k);
                gzp_co_save(V, k);
                return CO_R_BLOCKED;
            }
            co.evaluated = 0;                   // a shared node, a finalised one or a forced move: straight on
            break;
        }
        case CO_AFTER_EVAL: {
            // back in createNode / expandChild with the node in hand (evaluator.cpp:207-214, 232-236)
            const int N = co.node;
            if (co.evaluated) { c->num_evaluations++; c->total_evals++; }
            const int parent = co.req_parent, child = co.req_child;
            CHILD(c_to, NODE(n_child0, parent) + child) = N;
            gzp_child_refresh(V, parent, child, N);
            co.req_parent = -1;
            co.req_child = -1;
            if (gzp_is_finalised(V, N) || NODE(n_nchild, N) > 1) {
                gzp_path_push(V, N, -1);
                gzp_finish_playout(V, V.E->nodes[V.nb + N]);
                if (co.kind == CO_KIND_ONCE) { co.st = CO_DEAD; S.worker_count = gzp_sync_load(&S.worker_count) - 1; return CO_R_DONE; }
                co.st = gzp_co_loop_top(k);
            } else {
                NODE(n_inflight, N) = NODE(n_inflight, N) + 1;
                c->cur = N;
                co.st = CO_DESCEND;
            }
            break;
        }
        case CO_BAL_START: {
            // balanceFirstMoves (evaluator.cpp:282-339): children by traversals (std::sort, sortedTraversals :265-278)
            const int root = c->root;
            if (gzp_is_terminal(V, root)) { co.st = CO_DEAD; return CO_R_DONE; }
            const int nc = NODE(n_nchild, root);
            int mm = S.bal_max < nc ? S.bal_max : nc;
            if (mm > 32) mm = 32;
            gzp_cache_fill(V, root);
            ChildCache& C = V.C;
            int* order = C.order;
            for (int i = 0; i < nc; i++) order[i] = i;
            gzp_std_sort(order, nc, [&C](int a, int b) { return C.trav[a] > C.trav[b]; });
            GZP_SYNC();
            for (int i = 0; i < mm; i++) S.bal_order[i] = order[i];
            S.bal_max = mm;
            S.bal_i = 0;
            S.bal_wanted = mm > 0 ? (int)C.trav[order[0]] : 0;
            GZP_SYNC();
            co.st = CO_BAL_NEXT;
            break;
        }
        case CO_BAL_NEXT: {
            const int i = gzp_sync_load(&S.bal_i) + 1;
            GZP_SYNC();
            S.bal_i = i;
            GZP_SYNC();
            if (i >= gzp_sync_load(&S.bal_max)) { co.st = CO_DEAD; return CO_R_DONE; }
            co.st = CO_BAL_LOOP;
            break;
        }
        case CO_BAL_LOOP: {
            const int root = c->root;
            const int child = S.bal_order[gzp_sync_load(&S.bal_i)];
            const int ci = NODE(n_child0, root) + child;
            const int to = CHILD(c_to, ci);
            if ((int)CHILD(c_trav, ci) >= gzp_sync_load(&S.bal_wanted) || to < 0 || gzp_is_terminal(V, to)) { co.st = CO_BAL_NEXT; break; }
            // batch_size playouts below this child, each its own coroutine, then wait for all of them
            const int B = gzp_conf(V).batch_size < E->EB ? gzp_conf(V).batch_size : E->EB;
            for (int w = 1; w <= B; w++) {
                S.co[w].st = CO_BAL_WORKER;
                S.co[w].kind = CO_KIND_ONCE;
                S.co[w].cur = child;            // (which root child; the worker turns it into its path)
                gzp_queue_push(S.run, &S.n_run, w);
            }
            GZP_SYNC();
            S.worker_count = B;
            GZP_SYNC();
            co.st = CO_BAL_WAIT;
            break;
        }
        case CO_BAL_WAIT: {
            if (gzp_sync_load(&S.worker_count) > 0) { gzp_co_save(V, k); return CO_R_YIELD; }
            co.st = CO_BAL_LOOP;
            break;
        }
        case CO_BAL_WORKER: {
            // path = [(root, child)]; treePlayout(child's node, path)
            const int root = c->root;
            const int child = co.cur;
            const int to = CHILD(c_to, NODE(n_child0, root) + child);
            c->path_len = 0;
            gzp_path_push(V, root, child);
            c->cur = to;
            co.st = CO_DESCEND;
            break;
        }
        default:
            return CO_R_DONE;
        }
    }
    c->error = GZP_ERR_PATH;
    return CO_R_DONE;
}

// NetworkScheduler::mainLoop for one game. Returns true when the engine step ends here for this game (a propnet
// request is posted, the batch waits for the network, or the step's budget is spent); false when the search is over.
GZP_RARE bool gzp_sched_run(GameView& V, bool nn_fresh) {
    const PuctEngine* E = V.E;
    GameCtl* c = V.c;
    GzpSched& S = E->sched[V.g];
    if (!gzp_sync_load(&S.active)) gzp_sched_start(V);
    if (gzp_sync_load(&S.flush_wait)) {
        if (!nn_fresh) return true;                         // the batch is still with the network
        // "once we return, we must handle the results before doing anything else": every requestor gets its
        // answer, in order, and joins the runnables; then the yielders (scheduler.cpp:176-197)
        const int nreq = gzp_sync_load(&S.n_req);
        for (int i = 0; i < nreq; i++) {
            const int k = S.req[i];
            E->x_need_eval[(int64_t)V.g * E->EB + i] = 0;
            gzp_reply(V, S.co[k].node, i, false);
            gzp_queue_push(S.run, &S.n_run, k);
        }
        const int nyld = gzp_sync_load(&S.n_yld);
        for (int i = 0; i < nyld; i++) gzp_queue_push(S.run, &S.n_run, S.yld[i]);
        GZP_SYNC();
        S.n_req = 0; S.n_yld = 0; S.flush_wait = 0;
        GZP_SYNC();
    }
    for (int iter = 0; iter < 4 * GZP_MAX_EB; iter++) {
        int k = gzp_sync_load(&S.current);
        if (k < 0) {
            const int nrun = gzp_sync_load(&S.n_run), nreq = gzp_sync_load(&S.n_req), nyld = gzp_sync_load(&S.n_yld);
            bool jump = false;
            if (nrun == 0) {
                if (nreq == 0) {
                    if (nyld > 0) {
                        for (int i = 0; i < nyld; i++) gzp_queue_push(S.run, &S.n_run, S.yld[i]);
                        GZP_SYNC();
                        S.n_yld = 0;
                        GZP_SYNC();
                        continue;
                    }
                    // nothing left to run: the search's coroutine has returned
                    GZP_SYNC();
                    S.active = 0;
                    GZP_SYNC();
                    V.pn = E->co_path_node + (int64_t)V.g * (E->EB + 1) * E->MAXD;
                    V.pc = E->co_path_choice + (int64_t)V.g * (E->EB + 1) * E->MAXD;
                    c->path_len = 0;
                    c->in_playout = 0;
                    if (gzp_sync_load(&S.balance)) {           // balanceFirstMoves returns to the caller, no move is chosen
                        GZP_SYNC();
                        S.balance = 0;
                        GZP_SYNC();
                        c->phase = PH_IDLE;
                        c->cmd_done = 1;
                        return true;
                    }
                    c->phase = PH_MOVE_DONE;
                    return false;
                }
                jump = true;
            }
            const int B = gzp_conf(V).batch_size < E->EB ? gzp_conf(V).batch_size : E->EB;
            if (!jump && nreq >= (B < 1 ? 1 : B)) jump = true;
            if (jump) {
                // hand the batch to the network: rows 0 .. nreq-1 of this game were staged by gzp_step_post
                GZP_SYNC();
                for (int i = 0; i < nreq; i++) E->x_need_eval[(int64_t)V.g * E->EB + i] = 1;
                S.flush_wait = 1;
                GZP_SYNC();
                return true;
            }
            k = S.run[0];
            GZP_SYNC();
            for (int i = 1; i < nrun; i++) { const int v = S.run[i]; GZP_SYNC(); S.run[i - 1] = v; GZP_SYNC(); }
            S.n_run = nrun - 1;
            S.current = k;
            GZP_SYNC();
        }
        const int r = gzp_co_run(V, k);
        if (c->error) return true;
        if (r == CO_R_PROPNET) return true;                 // keeps the controller across the propnet pass
        GZP_SYNC();
        S.current = -1;
        GZP_SYNC();
        if (r == CO_R_YIELD) gzp_queue_push(S.yld, &S.n_yld, k);
    }
    return true;                                            // budget of this step spent; carries on next step
}

// Returns when the game is blocked on the pipeline (c->req != REQ_NONE or c->eval_node >= 0),
// idle, halted, or the per-step iteration budget is spent.
GZP_HDN void gzp_step_pre(GameView& V, bool nn_fresh = true) {
    const PuctEngine* E = V.E;
    GameCtl* c = V.c;
    if (c->phase == PH_HALT) return;
    if (E->EB > 1 && c->phase == PH_SEARCH) {
        // a Player's search with intra-tree workers: its own scheduler consumes the network's answers
        if (c->req != REQ_NONE) return;
        if (gzp_sched_run(V, nn_fresh)) return;
        nn_fresh = false;
    }
    if (c->eval_node >= 0) {
        if (!nn_fresh) return;          // still waiting for the network
        if (E->x_answered) {            // compacted batch: did this game's row fit into it?
            GZP_SYNC();
            const bool answered = E->x_answered[V.g] != 0;
            GZP_SYNC();
            if (!answered) return;
            E->x_answered[V.g] = 0;
        }
        E->x_need_eval[(int64_t)V.g * E->EB] = 0;
        gzp_reply(V, c->eval_node);
        c->num_evaluations++;
        c->total_evals++;
        c->eval_node = -1;
        c->resume = 1;
    }
    for (int iter = 0; iter < E->max_iters_per_step; iter++) {
        if (c->error) {
            if (gzp_restart_after_error(V)) continue;
            c->phase = PH_HALT;
            return;
        }
        if (V.work >= E->max_work_per_step) return;
        if (c->req != REQ_NONE) return;
        if (c->mode == 1 && c->phase == PH_AFTER_MOVE && !c->resume) {   // player: move applied
            c->phase = PH_IDLE;
            c->cmd_done = 1;
            continue;
        }
        if (c->resume) {
            c->resume = 0;
            const int N = c->new_node;
            if (c->req_purpose == PUR_PLAYOUT) {
                if (gzp_is_finalised(V, N) || NODE(n_nchild, N) > 1) {
                    gzp_path_push(V, N, -1);
                    gzp_finish_playout(V, V.E->nodes[V.nb + N]);
                } else {
                    NODE(n_inflight, N) = NODE(n_inflight, N) + 1;
                    c->cur = N;
                }
            } else if (c->req_purpose == PUR_REROOT) {
                gzp_request_reroot(V, N);
            } else {
                c->root = N;
                if (gzp_is_terminal(V, N)) { c->error = GZP_ERR_ROOT_TERMINAL; continue; }
                if (c->mode == 0) {
                    c->starting_sample_depth = NODE(n_depth, N);
                    c->conf_sel = 0;
                    c->phase = PH_COLLECT_NEXT;
                } else {
                    c->phase = PH_IDLE;
                    c->cmd_done = 1;
                }
            }
            continue;
        }
        switch (c->phase) {
        case PH_NEW_GAME:
            gzp_begin_game(V);
            break;
        case PH_COLLECT_NEXT: {
            const int node = c->root;
            c->conf_sel = 0;
            bool stop = false;
            if (E->conf.abort_max_length > 0 && NODE(n_depth, node) > E->conf.abort_max_length) stop = true;
            if (!stop && gzp_is_terminal(V, node)) stop = true;
            if (stop) { c->phase = PH_GAME_END; c->in_run_to_end = 0; break; }
            bool do_skip = false;
            if (!gzp_is_unique(E, NSTATE(node), NODE(n_depth, node), V.lane)) {
                c->saw_dupes++;
                do_skip = true;
            } else if (E->conf.oscillate_sampling_pct > 0 && gzp_rng_get(c->rng_sp) > E->conf.oscillate_sampling_pct) {
                do_skip = true;
            }
            if (!do_skip) {
                gzp_uniq_add(E, NSTATE(node), V.lane);
                gzp_reset_root_node(V);
                c->do_sample = 1;
                gzp_start_search(V, E->conf.evals_per_move, 0);
            } else {
                int skip_evals = (int)gzp_rng_with_max(c->rng_sp, (uint32_t)(E->conf.evals_per_move / 3 + 1));
                if (skip_evals < 16) skip_evals = 16;
                c->do_sample = 0;
                gzp_start_search(V, skip_evals, 1);
            }
            break;
        }
        case PH_SEARCH: {
            if (E->EB > 1) return;                     // a Player's search with intra-tree workers: the next step's
                                                       // gzp_sched_run (top of this function) takes it from here
            if (!c->in_playout) {
                const int root = c->root;
                const GzpNode root_nd = V.E->nodes[V.nb + root];
                GzpKid root_kid;
                bool have_kid = false;
                if (gzp_search_over(V, root_nd, &root_kid, &have_kid) || c->max_evaluations == 0) { c->phase = PH_MOVE_DONE; break; }
                c->cur = root;
                c->path_len = 0;
                c->in_playout = 1;
                if (gzp_playout_continue(V, &root_nd, have_kid ? &root_kid : nullptr)) return;
                break;
            }
            if (gzp_playout_continue(V)) return;
            break;
        }
        case PH_MOVE_DONE: {
            const int root = c->root;
            const int choice = gzp_choose(V, root);
            c->choice = choice;
            if (c->mode == 1) {       // player: search result is ready, wait for apply
                c->phase = PH_IDLE;
                c->cmd_done = 1;
                break;
            }
            if (!c->in_run_to_end) {
                if (c->do_sample) {
                    gzp_get_probabilities(V, root, E->conf.temperature_for_policy, false);
                    gzp_create_sample(V, root);
                } else {
                    c->conf_sel = 0;
                }
            }
            c->total_moves++;
            const int ci = NODE(n_child0, root) + choice;
            if (CHILD(c_to, ci) < 0) {
                gzp_request_expand(V, root, choice, PUR_REROOT);
                c->phase = PH_AFTER_MOVE;   // entered once the re-root completes
                return;
            }
            gzp_request_reroot(V, CHILD(c_to, ci));
            return;
        }
        case PH_AFTER_MOVE: {
            const int node = c->root;
            if (!c->in_run_to_end) {
                if (gzp_is_terminal(V, node)) { c->phase = PH_GAME_END; break; }
                if (!c->has_resigned) gzp_resign(V, node);
                if (c->has_resigned && c->n_samples > 1) {
                    c->number_resigns++;
                    c->phase = PH_GAME_END;
                    break;
                }
                c->phase = PH_COLLECT_NEXT;
            } else {
                if (gzp_is_finalised(V, node)) { c->phase = PH_GAME_END; break; }
                if (c->run_to_end_can_resign && NODE(n_depth, node) > E->conf.run_to_end_minimum_game_depth) {
                    const int lead = NODE(n_lead, node);
                    const float lead_score = NSCORE(node, lead);
                    if (lead_score < E->conf.run_to_end_early_score) {
                        c->number_early_run_to_ends++;
                        float fs[GZP_MAX_ROLES];
                        for (int r = 0; r < E->R; r++) fs[r] = (r == lead) ? 0.0f : 1.0f;
                        gzp_add_samples(V, fs, NODE(n_depth, node));
                        c->total_games++;
                        c->phase = PH_NEW_GAME;
                        break;
                    }
                }
                c->phase = PH_RUNTOEND_NEXT;
            }
            break;
        }
        case PH_GAME_END: {
            if (!c->in_run_to_end) {
                // end of collectSamples: either nothing sampled, or switch to runToEnd
                if (c->n_samples == 0) {
                    c->no_samples_taken++;
                    c->total_games++;
                    c->phase = PH_NEW_GAME;
                    break;
                }
                c->in_run_to_end = 1;
                c->conf_sel = 1;
                c->run_to_end_can_resign = (c->has_resigned && gzp_rng_get(c->rng_sp) > E->conf.run_to_end_pct) ? 1 : 0;
                c->phase = PH_RUNTOEND_NEXT;
                break;
            }
            // end of runToEnd
            const int node = c->root;
            if (E->conf.abort_max_length > 0 && NODE(n_depth, node) > E->conf.abort_max_length) {
                c->number_aborts_game_length++;
                c->n_samples = 0;
            } else {
                float fs[GZP_MAX_ROLES];
                for (int r = 0; r < E->R; r++) fs[r] = NSCORE(node, r);
                gzp_add_samples(V, fs, NODE(n_depth, node));
            }
            c->total_games++;
            c->phase = PH_NEW_GAME;
            break;
        }
        case PH_RUNTOEND_NEXT: {
            const int node = c->root;
            const bool done = (E->conf.abort_max_length > 0 && NODE(n_depth, node) > E->conf.abort_max_length) ||
                              gzp_is_finalised(V, node);
            if (done) { c->phase = PH_GAME_END; break; }
            c->do_sample = 0;
            gzp_start_search(V, E->conf.run_to_end_evals, 1);
            break;
        }
        case PH_IDLE: {
            if (c->cmd == CMD_NONE) return;
            const int cmd = c->cmd;
            c->cmd = CMD_NONE;
            if (cmd == CMD_RESET) {
                c->n_nodes = 0; c->n_children = 0; c->root = -1; c->root_has_prev = 0;
                gzp_tt_clear(V);
                c->game_depth = c->cmd_arg;
                c->in_run_to_end = 0; c->in_playout = 0; c->conf_sel = 0;
                // root state was placed in x_in_state by the host
                c->req = REQ_ROOT; c->req_parent = -1; c->req_child = -1; c->req_purpose = PUR_ROOT;
                for (int r = 0; r < E->R; r++) E->x_moves[(int64_t)V.g * E->R + r] = -1;
                return;
            } else if (cmd == CMD_APPLY) {
                // cmd_arg = lead-role action index of the move played
                const int root = c->root;
                int found = -1;
                for (int i = 0; i < NODE(n_nchild, root); i++)
                    if (CHILD(c_move, NODE(n_child0, root) + i) == c->cmd_arg) found = i;
                if (found < 0) { c->cmd_done = -1; break; }
                const int ci = NODE(n_child0, root) + found;
                if (CHILD(c_to, ci) < 0) {
                    gzp_request_expand(V, root, found, PUR_REROOT);
                    c->phase = PH_AFTER_MOVE;
                    return;
                }
                gzp_request_reroot(V, CHILD(c_to, ci));
                return;
            } else if (cmd == CMD_SEARCH) {
                gzp_start_search(V, c->cmd_evals, 0);
            } else if (cmd == CMD_BALANCE) {
                // Player::balanceNode (player.cpp:121-135): only a search tree with workers can do it here; without
                // them (batch_size 1) the call is accepted and does nothing
                if (E->EB > 1 && c->root >= 0 && c->cmd_arg > 0) {
                    GzpSched& S = E->sched[V.g];
                    GZP_SYNC();
                    S.balance = 1;
                    S.bal_max = c->cmd_arg;
                    S.active = 0;
                    GZP_SYNC();
                    c->conf_sel = 0;
                    c->in_playout = 0;
                    c->stop_request = GZP_STOP_NONE;
                    c->phase = PH_SEARCH;
                } else {
                    c->cmd_done = 1;
                }
            }
            break;
        }
        default:
            return;
        }
    }
}

// After the propnet pass: build the requested node from the scratch row and decide whether the
// network has to look at it.
GZP_HDN void gzp_step_post(GameView& V) {
    const PuctEngine* E = V.E;
    GameCtl* c = V.c;
    if (c->phase == PH_HALT || c->req == REQ_NONE || c->req == REQ_REROOT) return;
    const int parent = c->req == REQ_ROOT ? -1 : c->req_parent;
    const int depth = parent < 0 ? c->game_depth : NODE(n_depth, parent) + 1;
    // a Player's search with intra-tree workers: the expanding coroutine keeps the controller (gzp_sched_run) and
    // attaches the node to its parent's record itself, when the reference's createNode would have returned
    const bool workers = E->EB > 1 && c->phase == PH_SEARCH;
    if (parent >= 0 && E->tt_cap && gzp_conf(V).lookup_transpositions) {
        // expandChild (evaluator.cpp:216-239): the state is already in the tree at this depth -> share its node;
        // the playout then ends (or walks on through a forced move) without a new evaluation
        const int found = gzp_tt_find(V, E->x_state + (int64_t)V.g * E->SW, depth);
        if (found >= 0) {
            c->req = REQ_NONE;
            c->new_node = found;
            c->total_transpositions++;
            if (workers) return;                 // the coroutine attaches it when it resumes (gzp_co_run)
            CHILD(c_to, NODE(n_child0, parent) + c->req_child) = found;
            gzp_child_refresh(V, parent, c->req_child, found);
            c->resume = 1;
            return;
        }
    }
    const int N = gzp_create_node(V, parent, depth);
    c->req = REQ_NONE;
    if (N < 0) {
        if (!gzp_restart_after_error(V)) c->phase = PH_HALT;
        return;
    }
    if (parent >= 0 && !workers) {
        CHILD(c_to, NODE(n_child0, parent) + c->req_child) = N;
        gzp_child_refresh(V, parent, c->req_child, N);
    }
    c->new_node = N;
    if (gzp_conf(V).lookup_transpositions) gzp_tt_insert(E, V.g, V.nb, N, false);
    c->total_new_children += (uint64_t)NODE(n_nchild, N);
    const bool needs_eval = !gzp_is_finalised(V, N) && NODE(n_nchild, N) > 1;
    if (needs_eval) {
        c->eval_node = N;
        // the evaluation row: row 0 of the game, or — workers — the next free row of the batch being collected
        // (its request flag is raised when the batch goes to the network, gzp_sched_run)
        const int k = workers ? E->sched[V.g].n_req : 0;
        const int64_t row = (int64_t)V.g * E->EB + k;
        if (!workers) E->x_need_eval[row] = 1;
        E->e_prevcnt[row] = parent >= 0 ? 1 : (c->root_has_prev ? 1 : 0);
        if (E->EB > 1) {
            const uint32_t* xs = E->x_state + (int64_t)V.g * E->SW;
            const uint32_t* xp = E->x_in_state + (int64_t)V.g * E->SW;
            for (int w = 0; w < E->SW; w++) { E->e_state[row * E->SW + w] = xs[w]; E->e_prev[row * E->SW + w] = xp[w]; }
        }
    } else if (!workers) {
        c->resume = 1;
    }
}

#endif  // GZ_PUCT_CORE_H
