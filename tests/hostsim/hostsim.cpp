/* TEST INFRASTRUCTURE — CPU validation build of the PUCT core.
 *
 * Compiles galvanise_zero_b200/csrc/gz_puct_core.h (the very source the CUDA kernels run) with g++
 * and drives it with the CPU oracles in place of the device pipeline: oracle/propnet_ref.c for the
 * transition, oracle/encoder_ref.c for the planes and include/gz_synth.h for the network. Used by
 * tests/test_puct_hostsim.py to compare visit counts with the reference's own C++ (oracle/_ref)
 * without a GPU. Never loaded by the product path.
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../galvanise_zero_b200/csrc/gz_puct_layout.h"
#include "../../include/gz_synth.h"

extern "C" {
struct pnref_sm;
pnref_sm* pnref_create(const int32_t* blob, int64_t n_words);
void pnref_destroy(pnref_sm* sm);
int pnref_n_bases(const pnref_sm* sm);
int pnref_n_roles(const pnref_sm* sm);
int pnref_role_size(const pnref_sm* sm, int r);
int pnref_state_words(const pnref_sm* sm);
int pnref_legal_words(const pnref_sm* sm);
void pnref_initial_state(const pnref_sm* sm, uint32_t* out);
void pnref_eval_batch(pnref_sm* sm, int64_t n, const uint32_t* states, uint32_t* legal, uint8_t* terminal, uint8_t* goals);
void pnref_transition_batch(pnref_sm* sm, int64_t n, const uint32_t* states, const int16_t* moves,
                            uint32_t* next, uint32_t* legal, uint8_t* terminal, uint8_t* goals);
struct encref;
encref* encref_create(int channel_size, int channels_per_state, int num_control_channels, int num_prev_states);
void encref_add_board_base(encref* t, int base_indx, int buf_incr);
void encref_add_control_base(encref* t, int base_indx, int channel_id, float value);
int encref_total_size(const encref* t);
void encref_to_channels(const encref* t, const uint32_t* state, const uint32_t* const* prev_states, int n_prev, float* buf);
}

struct HostSim {
    PuctEngine E;
    std::vector<uint8_t> arena;
    pnref_sm* sm;
    encref* enc;
    std::vector<float> policy, value, planes;
    std::vector<uint8_t> answered;     // compacted batches: which games' predictions arrived with this step
    int answer_cap = 0;                // > 0: at most this many requests are answered per step (the rest wait)
    uint64_t nn_seed;
    int total;
    std::vector<int> legal_word_off;
};

static void fill_conf(PuctConf& p, const double* c) {
    int i = 0;
    p.puct_constant = (float)c[i++]; p.puct_constant_root = (float)c[i++];
    p.dirichlet_noise_pct = (float)c[i++]; p.noise_policy_squash_pct = (float)c[i++];
    p.noise_policy_squash_prob = (float)c[i++]; p.choose = c[i++] > 0.5 ? 1 : 0;
    p.random_scale = (float)c[i++]; p.temperature = (float)c[i++];
    p.depth_temperature_start = (int)c[i++]; p.depth_temperature_increment = (float)c[i++];
    p.depth_temperature_stop = (int)c[i++]; p.depth_temperature_max = (float)c[i++];
    p.fpu_prior_discount = (float)c[i++]; p.fpu_prior_discount_root = (float)c[i++];
    p.top_visits_best_guess_converge_ratio = (float)c[i++]; p.think_time = (float)c[i++];
    p.converged_visits = (int)c[i++]; p.batch_size = (int)c[i++]; p.use_legals_count_draw = (int)c[i++];
    p.backup_finalised = c[i++] > 0.5; p.lookup_transpositions = c[i++] > 0.5;
    p.evaluation_multiplier_to_convergence = (float)c[i++];
}

static void fill_sp(SpConf& s, const double* c) {
    int i = 0;
    s.oscillate_sampling_pct = (float)c[i++]; s.temperature_for_policy = (float)c[i++];
    s.evals_per_move = (int)c[i++]; s.resign0_score_probability = (float)c[i++]; s.resign0_pct = (float)c[i++];
    s.resign1_score_probability = (float)c[i++]; s.resign1_pct = (float)c[i++]; s.run_to_end_pct = (float)c[i++];
    s.run_to_end_evals = (int)c[i++]; s.run_to_end_early_score = (float)c[i++];
    s.run_to_end_minimum_game_depth = (int)c[i++]; s.abort_max_length = (int)c[i++];
}

extern "C" void* hs_create(const int32_t* blob, int64_t n_words, int G, int mode, int ncap, int ccap,
                           int channel_size, int channels_per_state, int num_control_channels,
                           int num_prev_states, int num_rewards, const int* board, int n_board,
                           const int* ctrl_bc, const float* ctrl_v, int n_ctrl, const double* sp_conf,
                           const double* puct_conf, const double* rte_conf, uint64_t nn_seed, uint64_t rng_seed) {
    HostSim* H = new HostSim();
    H->sm = pnref_create(blob, n_words);
    if (!H->sm) return nullptr;
    PuctEngine& E = H->E;
    memset(&E, 0, sizeof(E));
    E.G = G;
    E.R = pnref_n_roles(H->sm);
    E.SW = pnref_state_words(H->sm);
    E.LW = pnref_legal_words(H->sm);
    E.NCAP = ncap; E.CCAP = ccap; E.MAXD = 512;
    E.n_rewards = num_rewards;
    int off = 0, loff = 0, maxa = 0;
    for (int r = 0; r < E.R; r++) {
        E.role_sizes[r] = pnref_role_size(H->sm, r);
        E.policy_off[r] = off; off += E.role_sizes[r];
        E.legal_word_off[r] = loff; loff += (E.role_sizes[r] + 31) / 32;
        if (E.role_sizes[r] > maxa) maxa = E.role_sizes[r];
    }
    E.legal_word_off[E.R] = loff;
    E.policy_stride = off;
    E.max_actions = maxa;
    E.MAXPEND = 256;
    E.srec = gzp_sample_record_words(E.SW, maxa);
    E.READYCAP = 1 << 16;
    E.uniq_cap = 1 << 16;
    E.max_iters_per_step = 64;
    E.max_work_per_step = 1 << 20;
    E.prefetch_nodes = 0;
    E.dbg_cycles = nullptr;
    E.x_answered = nullptr;
    fill_sp(E.conf, sp_conf);
    fill_conf(E.conf.puct, puct_conf);
    fill_conf(E.conf.run_to_end, rte_conf);
    E.EB = 1;
    if (mode == 1 && E.conf.puct.batch_size > 1) E.EB = E.conf.puct.batch_size > GZP_MAX_EB ? GZP_MAX_EB : E.conf.puct.batch_size;
    E.tt_cap = 0;
    if (E.conf.puct.lookup_transpositions || E.conf.run_to_end.lookup_transpositions) {
        E.tt_cap = 32;
        while (E.tt_cap < 2 * E.NCAP) E.tt_cap *= 2;
    }
    const size_t bytes = gzp_layout(&E, nullptr);
    H->arena.assign(bytes, 0);
    gzp_layout(&E, H->arena.data());
    E.seed = rng_seed;
    // init state and hash mask (every base the transformer looks at, gdltransformer.cpp:54-63)
    pnref_initial_state(H->sm, (uint32_t*)E.init_state);
    uint32_t* mask = (uint32_t*)E.uniq_mask;
    H->enc = encref_create(channel_size, channels_per_state, num_control_channels, num_prev_states);
    for (int i = 0; i < n_board; i++) {
        encref_add_board_base(H->enc, board[2 * i], board[2 * i + 1]);
        mask[board[2 * i] >> 5] |= 1u << (board[2 * i] & 31);
    }
    for (int i = 0; i < n_ctrl; i++) {
        encref_add_control_base(H->enc, ctrl_bc[2 * i], ctrl_bc[2 * i + 1], ctrl_v[i]);
        mask[ctrl_bc[2 * i] >> 5] |= 1u << (ctrl_bc[2 * i] & 31);
    }
    H->total = encref_total_size(H->enc);
    H->policy.assign((size_t)G * E.EB * E.policy_stride, 0.0f);
    H->value.assign((size_t)G * E.EB * num_rewards, 0.0f);
    H->planes.assign((size_t)H->total, 0.0f);
    E.nn_policy = H->policy.data();
    E.nn_value = H->value.data();
    H->nn_seed = nn_seed;
    // per-game rng streams: evaluator then selfplay, in construction order (as the shim does)
    uint64_t sd = rng_seed;
    for (int g = 0; g < G; g++) {
        GameCtl& c = E.ctl[g];
        c.mode = mode;
        c.phase = mode == 0 ? PH_NEW_GAME : PH_IDLE;
        c.eval_node = -1;
        gzp_rng_seed(c.rng_eval, sd); sd += 0x9E3779B97F4A7C15ULL;
        gzp_rng_seed(c.rng_sp, sd); sd += 0x9E3779B97F4A7C15ULL;
    }
    return H;
}

extern "C" void hs_destroy(void* h) {
    HostSim* H = (HostSim*)h;
    pnref_destroy(H->sm);
    delete H;
}

/* one full pipeline step for every game: pre -> propnet -> post -> encode -> network */
extern "C" void hs_step(void* h) {
    HostSim* H = (HostSim*)h;
    PuctEngine& E = H->E;
    for (int g = 0; g < E.G; g++) {
        GameView V = gzp_view(&E, g);
        gzp_step_pre(V);
    }
    for (int g = 0; g < E.G; g++) {      // the device runs puct_reroot_kernel here
        GameCtl& c = E.ctl[g];
        if (c.req != REQ_REROOT) continue;
        GameView V = gzp_view(&E, g);
        gzp_reroot_host(V, c.req_parent);
        c.req = REQ_NONE;
    }
    for (int g = 0; g < E.G; g++) {
        GameCtl& c = E.ctl[g];
        if (c.req == REQ_NONE) continue;
        const uint32_t* in = E.x_in_state + (size_t)g * E.SW;
        uint32_t* out = E.x_state + (size_t)g * E.SW;
        if (E.x_moves[(size_t)g * E.R] < 0) {
            memcpy(out, in, sizeof(uint32_t) * E.SW);
            pnref_eval_batch(H->sm, 1, in, E.x_legal + (size_t)g * E.LW, E.x_term + g, E.x_goals + (size_t)g * E.R);
        } else {
            pnref_transition_batch(H->sm, 1, in, E.x_moves + (size_t)g * E.R, out, E.x_legal + (size_t)g * E.LW,
                                   E.x_term + g, E.x_goals + (size_t)g * E.R);
        }
    }
    for (int g = 0; g < E.G; g++) {
        GameView V = gzp_view(&E, g);
        gzp_step_post(V);
    }
    if (getenv("GZREF_TRACE")) { int n = 0; for (int g = 0; g < E.G * E.EB; g++) n += E.x_need_eval[g] != 0; if (n) fprintf(stderr, "B %d\n", n); }
    int answered_now = 0;
    for (int g = 0; g < E.G * E.EB; g++) {          // evaluation rows
        if (!E.x_need_eval[g]) continue;
        if (H->answer_cap > 0) {                     // gz_puct_compact_requests: ascending rows, capped; overflow waits
            if (answered_now >= H->answer_cap) continue;
            H->answered[g] = 1;
            answered_now++;
        }
        const uint32_t* prev = E.e_prev + (size_t)g * E.SW;
        const uint32_t* ptrs[1] = {prev};
        encref_to_channels(H->enc, E.e_state + (size_t)g * E.SW, ptrs, E.e_prevcnt[g] ? 1 : 0, H->planes.data());
        const uint64_t hh = gz_synth_hash_planes_f32(H->planes.data(), H->total, H->nn_seed);
        for (int r = 0; r < E.R; r++)
            for (int a = 0; a < E.role_sizes[r]; a++)
                H->policy[(size_t)g * E.policy_stride + E.policy_off[r] + a] = gz_synth_policy(hh, (uint32_t)r, (uint32_t)a);
        gz_synth_values(hh, E.n_rewards, &H->value[(size_t)g * E.n_rewards]);
    }
}

/* compacted batches (gz_puct_compact_requests / gz_puct_advance_compact): at most `cap` requests per step get their
 * prediction, the others stay pending */
extern "C" void hs_set_answer_cap(void* h, int cap) {
    HostSim* H = (HostSim*)h;
    H->answer_cap = cap;
    H->answered.assign((size_t)H->E.G * H->E.EB, 0);
    H->E.x_answered = cap > 0 ? H->answered.data() : nullptr;
}

extern "C" GameCtl* hs_ctl(void* h, int g) { return &((HostSim*)h)->E.ctl[g]; }

/* player-mode helpers */
extern "C" void hs_cmd_reset(void* h, int g, const uint32_t* state, int game_depth) {
    HostSim* H = (HostSim*)h;
    PuctEngine& E = H->E;
    const uint32_t* s = state ? state : E.init_state;
    memcpy(E.x_in_state + (size_t)g * E.SW, s, sizeof(uint32_t) * E.SW);
    E.ctl[g].cmd = CMD_RESET; E.ctl[g].cmd_arg = game_depth; E.ctl[g].cmd_done = 0;
}
extern "C" void hs_cmd_search(void* h, int g, int evals) {
    GameCtl& c = ((HostSim*)h)->E.ctl[g];
    c.cmd = CMD_SEARCH; c.cmd_evals = evals; c.cmd_done = 0;
}
extern "C" void hs_cmd_balance(void* h, int g, int max_moves) {
    GameCtl& c = ((HostSim*)h)->E.ctl[g];
    c.cmd = CMD_BALANCE; c.cmd_arg = max_moves; c.cmd_done = 0;
}
extern "C" void hs_cmd_apply(void* h, int g, int lead_move) {
    GameCtl& c = ((HostSim*)h)->E.ctl[g];
    c.cmd = CMD_APPLY; c.cmd_arg = lead_move; c.cmd_done = 0;
}
extern "C" int hs_wait(void* h, int g, int max_steps) {
    HostSim* H = (HostSim*)h;
    for (int i = 0; i < max_steps; i++) {
        if (H->E.ctl[g].cmd_done) return H->E.ctl[g].cmd_done;
        if (H->E.ctl[g].phase == PH_HALT) return -100 - H->E.ctl[g].error;
        hs_step(h);
    }
    return 0;
}
/* root children statistics of game g; returns num_children */
extern "C" int hs_root_children(void* h, int g, int max_children, int* moves, int* trav, float* prob, float* score,
                                float* root_info) {
    HostSim* H = (HostSim*)h;
    GameView V = gzp_view(&H->E, g);
    const int n = V.c->root;
    const int nc = NODE(n_nchild, n), c0 = NODE(n_child0, n);
    const int lead = NODE(n_lead, n);
    for (int i = 0; i < nc && i < max_children; i++) {
        moves[i] = CHILD(c_move, c0 + i);
        trav[i] = (int)CHILD(c_trav, c0 + i);
        prob[i] = CHILD(c_prob, c0 + i);
        const int to = CHILD(c_to, c0 + i);
        score[i] = to >= 0 ? NSCORE(to, lead) : -1.0f;
    }
    root_info[0] = (float)NODE(n_visits, n);
    root_info[1] = (float)lead;
    root_info[2] = NSCORE(n, lead);
    root_info[3] = (float)V.c->n_nodes;
    root_info[4] = (float)V.c->choice;
    return nc;
}

extern "C" unsigned long long hs_ready_count(void* h) { return *((HostSim*)h)->E.ready_count; }
extern "C" const int32_t* hs_ready(void* h) { return ((HostSim*)h)->E.ready; }
extern "C" unsigned long long hs_transpositions(void* h, int g) { return ((HostSim*)h)->E.ctl[g].total_transpositions; }
extern "C" int hs_srec(void* h) { return ((HostSim*)h)->E.srec; }
extern "C" int hs_sizeof_ctl() { return (int)sizeof(GameCtl); }
