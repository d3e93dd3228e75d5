// gz_host.cu — the host side of the reference's Supervisor and Player, native, over the device engine.
//
// Reference (compiled C++ behind the extension module):
//   Supervisor   src/cpp/supervisor.cpp:20-245   createInline / createWorkers / poll / getSamples
//   Player       src/cpp/player.cpp:20-173       puctPlayerReset / ApplyMove / Move / GetMove / poll
//   bindings     src/cpp/pyobjects/{supervisor,player}_impl.cpp, common.cpp:160-218 (the poll contract)
//
// gz_supervisor: `start_self_play(num_workers <= 0)` is the reference's inline mode (one manager of batch_size
// games, strictly serial); `num_workers > 0` its worker mode (supervisor.cpp:79-99: two managers per worker so
// that the search of one overlaps the network call of the other) — here two engines of batch_size games on
// their own CUDA streams, handed out alternately as ready managers are in the reference (supervisor.cpp:129-167).
// poll() takes the predictions for the rows handed out by the previous poll (host float32 arrays, one per
// policy head + values) and returns the planes of the next batch in pinned host memory owned by the engine.
//
// gz_player: player_apply_move / player_move only enqueue work (player.cpp:59-103); poll() carries it out, hands
// out the planes of the one leaf to evaluate, and returns 0 once the queue has drained. The wall-clock limits
// of playoutMain (evaluator.cpp:794,820-826) are watched here on the host and passed to the device search
// as stop requests (gz_puct_request_stop).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <deque>
#include <string>
#include <vector>

#include "../../include/gz_b200.h"
#include "gz_common.cuh"

static double gz_now() {
    return std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------------------------------------
struct gz_supervisor {
    gz_propnet* pn = nullptr;
    gz_transformer* tr = nullptr;
    int batch_size = 0, device = 0, node_cap = 4096, child_cap = 0;
    uint64_t seed = 0;
    std::string identifier;
    struct Manager {
        gz_puct* eng = nullptr;
        cudaStream_t stream = nullptr;
        int n_req = 0;
        const float* planes = nullptr;
    };
    std::vector<Manager> man;
    int cur = -1;
    bool primed = false;
    int n_roles = 0;
};

extern "C" int gz_supervisor_create(gz_propnet* pn, gz_transformer* tr, int batch_size, const char* identifier, int device,
                                    uint64_t seed, int node_cap, int child_cap, gz_supervisor** out) {
    if (!pn || !tr || !out || batch_size <= 0) { gz_set_error("bad argument"); return GZ_ERR_ARG; }
    gz_supervisor* s = new gz_supervisor();
    s->pn = pn; s->tr = tr; s->batch_size = batch_size; s->device = device; s->seed = seed;
    s->identifier = identifier ? identifier : "";
    s->node_cap = node_cap > 0 ? node_cap : 4096;
    s->n_roles = (int)gz_propnet_info(pn, GZ_INFO_N_ROLES);
    int maxa = 0;
    for (int r = 0; r < s->n_roles; r++) maxa = std::max(maxa, (int)gz_propnet_info(pn, GZ_INFO_ROLE_SIZE0 + r));
    s->child_cap = child_cap > 0 ? child_cap : s->node_cap * std::max(8, std::min(maxa, 64) / 2);
    *out = s;
    return GZ_OK;
}

static void supervisor_drop_managers(gz_supervisor* s) {
    for (auto& m : s->man) {
        if (m.eng) gz_puct_destroy(m.eng);
        if (m.stream) cudaStreamDestroy(m.stream);
    }
    s->man.clear();
}

extern "C" void gz_supervisor_destroy(gz_supervisor* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    supervisor_drop_managers(s);
    delete s;
}

extern "C" int gz_supervisor_start_self_play(gz_supervisor* s, int num_workers, const double* sp_conf,
                                             const double* puct_conf, const double* rte_conf) {
    if (!s || !sp_conf || !puct_conf || !rte_conf) { gz_set_error("null argument"); return GZ_ERR_ARG; }
    GZ_CUDA(cudaSetDevice(s->device));
    supervisor_drop_managers(s);
    const int n_man = num_workers > 0 ? 2 : 1;
    for (int k = 0; k < n_man; k++) {
        gz_supervisor::Manager m;
        int rc = gz_puct_create(s->pn, s->tr, s->batch_size, GZ_PUCT_MODE_SELFPLAY, s->node_cap, s->child_cap, 0, 0,
                                sp_conf, puct_conf, rte_conf, s->seed + 104729ull * (uint64_t)k, &m.eng);
        if (rc != GZ_OK) { supervisor_drop_managers(s); return rc; }
        GZ_CUDA(cudaStreamCreateWithFlags(&m.stream, cudaStreamNonBlocking));
        s->man.push_back(m);
    }
    GZ_CUDA(cudaDeviceSynchronize());
    s->cur = -1;
    s->primed = false;
    return GZ_OK;
}

// what: 0 managers, 1 batch size, 2 plane row size, 3 sample record words, 4 roles
extern "C" int64_t gz_supervisor_info(const gz_supervisor* s, int what) {
    if (!s) return -1;
    switch (what) {
        case 0: return (int64_t)s->man.size();
        case 1: return s->batch_size;
        case 2: return s->man.empty() ? -1 : gz_puct_info(s->man[0].eng, 1);
        case 3: return s->man.empty() ? -1 : gz_puct_info(s->man[0].eng, 4);
        case 4: return s->n_roles;
        default: return -1;
    }
}

extern "C" gz_puct* gz_supervisor_engine(gz_supervisor* s, int k) {
    return (s && k >= 0 && k < (int)s->man.size()) ? s->man[k].eng : nullptr;
}
extern "C" void* gz_supervisor_stream(gz_supervisor* s, int k) {
    return (s && k >= 0 && k < (int)s->man.size()) ? (void*)s->man[k].stream : nullptr;
}

// Supervisor::poll (supervisor.cpp:101-168, common.cpp:160-218). Returns the number of rows handed out (> 0,
// *h_planes = [rows][total] float32, valid until the same manager's batch is handed out again), 0 when there is
// nothing left to evaluate, or a negative gz_status.
extern "C" int gz_supervisor_poll(gz_supervisor* s, int n_pred, const float* const* h_arrays, int n_arrays,
                                  const float** h_planes) {
    if (!s || s->man.empty()) { gz_set_error("start_self_play() has not been called"); return GZ_ERR_ARG; }
    if (n_pred < 0 || n_pred > s->batch_size) { gz_set_error("n_predictions out of range"); return GZ_ERR_ARG; }
    GZ_CUDA(cudaSetDevice(s->device));
    const int M = (int)s->man.size();
    int nxt;
    int rc;
    if (!s->primed) {
        if (n_pred != 0) { gz_set_error("first poll must pass n_predictions == 0"); return GZ_ERR_ARG; }
        for (auto& m : s->man) {
            rc = gz_puct_host_poll_submit(m.eng, 0, nullptr, 0, (void*)m.stream);
            if (rc != GZ_OK) return rc;
        }
        s->primed = true;
        nxt = 0;
    } else {
        gz_supervisor::Manager& m = s->man[s->cur];
        if (n_pred != m.n_req) { gz_set_error("expected %d predictions, got %d", m.n_req, n_pred); return GZ_ERR_ARG; }
        m.n_req = 0;
        rc = gz_puct_host_poll_submit(m.eng, n_pred, n_pred ? h_arrays : nullptr, n_pred ? n_arrays : 0, (void*)m.stream);
        if (rc != GZ_OK) return rc;
        nxt = (s->cur + 1) % M;
    }
    // hand out the next ready batch; a manager with nothing to evaluate is stepped again
    for (int tries = 0; tries < 64 * M; tries++) {
        gz_supervisor::Manager& m = s->man[nxt];
        const int n = gz_puct_host_poll_wait(m.eng, &m.planes);
        if (n < 0) return n;
        if (n > 0) {
            m.n_req = n;
            s->cur = nxt;
            if (h_planes) *h_planes = m.planes;
            return n;
        }
        rc = gz_puct_host_poll_submit(m.eng, 0, nullptr, 0, (void*)m.stream);
        if (rc != GZ_OK) return rc;
        nxt = (nxt + 1) % M;
    }
    for (auto& m : s->man) {                // leave no poll in flight
        const int n = gz_puct_host_poll_wait(m.eng, &m.planes);
        if (n < 0) return n;
    }
    s->primed = false;
    s->cur = -1;
    return 0;
}

// Finished samples of all managers (records of gz_supervisor_info(s, 3) int32 words); h_manager[i] (optional) =
// the manager record i came from. Returns the number of records (<= cap).
extern "C" int64_t gz_supervisor_fetch_samples(gz_supervisor* s, int32_t* h_out, int64_t cap, int32_t* h_manager) {
    if (!s || (!h_out && cap > 0)) { gz_set_error("null argument"); return GZ_ERR_ARG; }
    cudaSetDevice(s->device);
    int64_t n = 0;
    for (size_t k = 0; k < s->man.size() && n < cap; k++) {
        const int64_t srec = gz_puct_info(s->man[k].eng, 4);
        const int64_t got = gz_puct_fetch_samples(s->man[k].eng, h_out + n * srec, cap - n, (void*)s->man[k].stream);
        if (got < 0) return got;
        if (h_manager) for (int64_t i = 0; i < got; i++) h_manager[n + i] = (int32_t)k;
        n += got;
    }
    return n;
}

extern "C" int gz_supervisor_clear_unique_states(gz_supervisor* s) {
    if (!s) return GZ_ERR_ARG;
    for (auto& m : s->man) {
        int rc = gz_puct_clear_unique_states(m.eng, (void*)m.stream);
        if (rc != GZ_OK) return rc;
    }
    return GZ_OK;
}

// Sums (max_* / first_error: maxima) of gz_puct_totals over the managers.
extern "C" int gz_supervisor_totals(gz_supervisor* s, double* h_out26) {
    if (!s || !h_out26) { gz_set_error("null argument"); return GZ_ERR_ARG; }
    for (int i = 0; i < 26; i++) h_out26[i] = 0;
    for (auto& m : s->man) {
        double t[26];
        int rc = gz_puct_totals(m.eng, t, (void*)m.stream);
        if (rc != GZ_OK) return rc;
        for (int i = 0; i < 26; i++) {
            if (i == 7 || i == 8 || i == 9) h_out26[i] = std::max(h_out26[i], t[i]);
            else h_out26[i] += t[i];
        }
    }
    return GZ_OK;
}

// ------------------------------------------------------------------------------------------------
struct gz_player {
    gz_propnet* pn = nullptr;
    gz_transformer* tr = nullptr;
    gz_puct* eng = nullptr;
    int device = 0, n_roles = 0, total = 0, policy_stride = 0, n_rewards = 0, sw = 0, max_actions = 0;
    std::vector<int> role_sizes, policy_off;
    cudaStream_t stream = nullptr;
    float *d_policy = nullptr, *d_value = nullptr, *d_planes = nullptr, *h_planes = nullptr, *h_row = nullptr;
    std::vector<double> sp_conf, puct_conf;
    struct Cmd { int kind; std::vector<int> moves; std::vector<uint32_t> state; int evals; double end_time; };
    std::deque<Cmd> queue;      // kind: 1 reset, 2 apply, 3 search, 4 balance the first `evals` root moves
    bool busy = false, fresh = false, first_play = false;
    int current = 0, waiting = 0, game_depth = 0, rows = 1;
    double search_start = 0, search_end_time = -1;
    int stop_sent = 0;
    // the finished search
    bool have_choice = false;
    int ch_i[16];
    float ch_f[8];
    std::vector<int> ch_moves, ch_trav;
    std::vector<float> ch_prob, ch_score, ch_next;
    int ch_n = 0;
};

enum { PC_THINK_TIME = 15, PC_CONVERGED_VISITS = 16, PC_EVAL_MULT = 21, PC_COUNT = 22 };   // puct conf array slots

extern "C" int gz_player_create(gz_propnet* pn, gz_transformer* tr, const double* puct_conf, int device, uint64_t seed,
                                int node_cap, int child_cap, gz_player** out) {
    if (!pn || !tr || !puct_conf || !out) { gz_set_error("null argument"); return GZ_ERR_ARG; }
    GZ_CUDA(cudaSetDevice(device));
    gz_player* p = new gz_player();
    p->pn = pn; p->tr = tr; p->device = device;
    p->n_roles = (int)gz_propnet_info(pn, GZ_INFO_N_ROLES);
    p->puct_conf.assign(puct_conf, puct_conf + PC_COUNT);
    // the self-play part of the configuration is unused in player mode (template values, templates.py:106-129)
    const double sp[12] = {0.25, 1.0, 100, 0.1, 0.99, 0.025, 0.95, 0.01, 32, 0.01, 30, -1};
    p->sp_conf.assign(sp, sp + 12);
    int maxa = 0, off = 0;
    for (int r = 0; r < p->n_roles; r++) {
        const int a = (int)gz_propnet_info(pn, GZ_INFO_ROLE_SIZE0 + r);
        p->role_sizes.push_back(a);
        p->policy_off.push_back(off);
        off += a;
        maxa = std::max(maxa, a);
    }
    if (node_cap <= 0) node_cap = 1 << 16;
    if (child_cap <= 0) child_cap = node_cap * std::max(8, std::min(maxa, 64) / 2);
    int rc = gz_puct_create(pn, tr, 1, GZ_PUCT_MODE_PLAYER, node_cap, child_cap, 0, 0, p->sp_conf.data(),
                            p->puct_conf.data(), p->puct_conf.data(), seed, &p->eng);
    if (rc != GZ_OK) { delete p; return rc; }
    p->total = (int)gz_puct_info(p->eng, 1);
    p->policy_stride = (int)gz_puct_info(p->eng, 2);
    p->n_rewards = (int)gz_puct_info(p->eng, 3);
    p->sw = (int)gz_puct_info(p->eng, 6);
    p->max_actions = (int)gz_puct_info(p->eng, 7);
    p->rows = (int)gz_puct_info(p->eng, 10);       // leaves one poll may hand out: PuctConfig.batch_size (player.cpp:32)
    const size_t EB = (size_t)p->rows;
    GZ_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    GZ_CUDA(cudaMalloc((void**)&p->d_policy, EB * p->policy_stride * 4));
    GZ_CUDA(cudaMalloc((void**)&p->d_value, EB * p->n_rewards * 4));
    GZ_CUDA(cudaMalloc((void**)&p->d_planes, EB * p->total * 4));
    GZ_CUDA(cudaMemset(p->d_policy, 0, EB * p->policy_stride * 4));
    GZ_CUDA(cudaMemset(p->d_value, 0, EB * p->n_rewards * 4));
    GZ_CUDA(cudaHostAlloc((void**)&p->h_planes, EB * p->total * 4, cudaHostAllocDefault));
    GZ_CUDA(cudaHostAlloc((void**)&p->h_row, EB * (p->policy_stride + p->n_rewards) * 4, cudaHostAllocDefault));
    p->ch_moves.resize(maxa); p->ch_trav.resize(maxa); p->ch_prob.resize(maxa); p->ch_score.resize(maxa); p->ch_next.resize(maxa);
    *out = p;
    return GZ_OK;
}

extern "C" void gz_player_destroy(gz_player* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->eng) gz_puct_destroy(p->eng);
    if (p->d_policy) cudaFree(p->d_policy);
    if (p->d_value) cudaFree(p->d_value);
    if (p->d_planes) cudaFree(p->d_planes);
    if (p->h_planes) cudaFreeHost(p->h_planes);
    if (p->h_row) cudaFreeHost(p->h_row);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

extern "C" gz_puct* gz_player_engine(gz_player* p) { return p ? p->eng : nullptr; }
extern "C" int64_t gz_player_info(const gz_player* p, int what) {   // 0 plane row size, 1 roles, 2 max actions, 3 state words, 4 most rows per poll
    if (!p) return -1;
    switch (what) { case 0: return p->total; case 1: return p->n_roles; case 2: return p->max_actions; case 3: return p->sw; case 4: return p->rows; default: return -1; }
}

// puctPlayerReset (player.cpp:52-57)
extern "C" int gz_player_reset(gz_player* p, int game_depth) {
    if (!p) return GZ_ERR_ARG;
    p->game_depth = game_depth;
    p->first_play = true;
    p->queue.clear();
    p->have_choice = false;
    return GZ_OK;
}

// puctPlayerUpdateConfig (player.cpp:40-50): think_time, converged_visits (verbose has no device counterpart)
extern "C" int gz_player_update_config(gz_player* p, double think_time, int converged_visits, int verbose) {
    if (!p) return GZ_ERR_ARG;
    (void)verbose;
    p->puct_conf[PC_THINK_TIME] = think_time;
    p->puct_conf[PC_CONVERGED_VISITS] = converged_visits;
    return gz_puct_set_config(p->eng, p->sp_conf.data(), p->puct_conf.data(), p->puct_conf.data());
}

// puctApplyMove (player.cpp:59-78): enqueued; the first call after a reset establishes the root first
extern "C" int gz_player_apply_move(gz_player* p, const int* joint_move, int n_roles) {
    if (!p || !joint_move || n_roles != p->n_roles) { gz_set_error("joint move must hold one index per role"); return GZ_ERR_ARG; }
    if (p->first_play) {
        p->first_play = false;
        p->queue.push_back({1, {}, {}, 0, -1});                    // establishRoot(nullptr): the initial state
    }
    p->queue.push_back({2, std::vector<int>(joint_move, joint_move + n_roles), {}, 0, -1});
    return GZ_OK;
}

// puctPlayerMove (player.cpp:80-103): enqueued; h_state = packed base state (state words) or null
extern "C" int gz_player_move(gz_player* p, const uint32_t* h_state, int evaluations, double end_time) {
    if (!p) return GZ_ERR_ARG;
    p->have_choice = false;
    if (p->first_play) {
        p->first_play = false;
        gz_player::Cmd c{1, {}, {}, 0, -1};
        if (h_state) c.state.assign(h_state, h_state + p->sw);
        p->queue.push_back(c);
    }
    p->queue.push_back({3, {}, {}, evaluations, end_time});
    return GZ_OK;
}

// puctPlayerGetMove (player.cpp:105-119): the chosen child's move for `role`, ITS score for `role`, node count
extern "C" int gz_player_get_move(gz_player* p, int role, int* move, float* score, int* nodes) {
    if (!p || role < 0 || role >= p->n_roles) { gz_set_error("bad role"); return GZ_ERR_ARG; }
    *move = -1; *score = -1.0f; *nodes = -1;
    if (!p->have_choice || p->ch_i[3] < 0 || p->ch_n == 0) return GZ_OK;
    const int ch = p->ch_i[3], lead = p->ch_i[7];
    *move = role == lead ? p->ch_moves[ch] : p->ch_i[9 + role];
    *score = p->ch_f[1 + role];
    *nodes = p->ch_i[4];
    return GZ_OK;
}

// Player::balanceNode (player.cpp:121-135) -> PuctEvaluator::balanceFirstMoves (evaluator.cpp:282-339): enqueued; the
// following polls search the first max_count root moves until each has as many traversals as the most visited one.
// Needs the search tree's workers (PuctConfig.batch_size > 1, the reference's default); with batch_size 1 the call
// is accepted and does nothing.
extern "C" int gz_player_balance_moves(gz_player* p, int max_count) {
    if (!p) return GZ_ERR_ARG;
    if (max_count > 0 && p->rows > 1 && !p->first_play) p->queue.push_back({4, {}, {}, max_count, -1});
    return GZ_OK;
}

// treeDebug (player.cpp:137-153, node.cpp:376-441): the root's expanded children by traversals (then score) with the
// principal variation below each (10 plies at most, nodes with at least 100 visits). scores / lead / moves: one per
// line; var_len[i] pairs (lead role, move) at var_pairs + i * 20. Returns the count written (<= max_count).
extern "C" int gz_puct_tree_debug(gz_puct* h, int g, int max_count, int depth, int* h_out_i, float* h_out_f, void* stream);
extern "C" int gz_player_tree_debug_lines(gz_player* p, int max_count, float* scores, int* lead, int* moves, int* var_len,
                                          int* var_pairs) {
    if (!p || !scores || !lead || !moves) { gz_set_error("null argument"); return GZ_ERR_ARG; }
    if (max_count <= 0) return 0;
    const int depth = 10, stride = 3 + 2 * depth;
    std::vector<int> oi((size_t)max_count * stride);
    std::vector<float> of(max_count);
    const int n = gz_puct_tree_debug(p->eng, 0, max_count, depth, oi.data(), of.data(), (void*)p->stream);
    if (n < 0) return n;
    for (int i = 0; i < n; i++) {
        const int* li = oi.data() + (size_t)i * stride;
        scores[i] = of[i]; lead[i] = li[0]; moves[i] = li[1];
        if (var_len) var_len[i] = li[2];
        if (var_pairs) for (int k = 0; k < 2 * li[2]; k++) var_pairs[i * 2 * depth + k] = li[3 + k];
    }
    return n;
}
extern "C" int gz_player_tree_debug(gz_player* p, int max_count, float* scores, int* lead, int* moves) {
    return gz_player_tree_debug_lines(p, max_count, scores, lead, moves, nullptr, nullptr);
}

static int player_issue_next(gz_player* p) {
    gz_player::Cmd c = p->queue.front();
    p->queue.pop_front();
    p->current = c.kind;
    int rc = GZ_OK;
    if (c.kind == 1) {
        rc = gz_puct_command(p->eng, 0, GZ_PUCT_CMD_RESET, p->game_depth, 0, c.state.empty() ? nullptr : c.state.data(), (void*)p->stream);
        if (rc == GZ_OK && !c.state.empty()) GZ_CUDA(cudaStreamSynchronize(p->stream));   // the host buffer is local
    } else if (c.kind == 2) {
        int oi[16];
        float of[8];
        const int n = gz_puct_snapshot(p->eng, 0, 0, oi, of, nullptr, nullptr, nullptr, nullptr, nullptr, (void*)p->stream);
        if (n < 0) return n;
        const int lead = oi[7] >= 0 ? oi[7] : 0;
        rc = gz_puct_command(p->eng, 0, GZ_PUCT_CMD_APPLY, c.moves[lead], 0, nullptr, (void*)p->stream);
        p->game_depth++;
    } else if (c.kind == 4) {
        rc = gz_puct_command(p->eng, 0, GZ_PUCT_CMD_BALANCE, c.evals, 0, nullptr, (void*)p->stream);
    } else {
        rc = gz_puct_command(p->eng, 0, GZ_PUCT_CMD_SEARCH, 0, c.evals, nullptr, (void*)p->stream);
        p->search_start = gz_now();
        p->search_end_time = c.end_time;
        p->stop_sent = 0;
    }
    p->busy = true;
    return rc;
}

// the clock of playoutMain (evaluator.cpp:794: hard end_time; :820-826: think_time, x multiplier when not converged)
static int player_watch_clock(gz_player* p) {
    if (p->current != 3) return GZ_OK;
    const double now = gz_now();
    const double think = p->puct_conf[PC_THINK_TIME];
    int want = 0;
    if (p->search_end_time > 0 && now > p->search_end_time) want = 1;
    else if (think > 0 && now > p->search_start + think * p->puct_conf[PC_EVAL_MULT]) want = 1;
    else if (think > 0 && now > p->search_start + think) want = 2;
    if (want && want != p->stop_sent && (p->stop_sent == 0 || want == 1)) {
        int rc = gz_puct_request_stop(p->eng, 0, want, (void*)p->stream);
        if (rc != GZ_OK) return rc;
        p->stop_sent = want;
    }
    return GZ_OK;
}

// Player::poll (player.cpp:155-173, common.cpp:160-218): returns n >= 1 with *h_planes = the n leaves to evaluate
// ([n][total] float32, pinned, valid until the next poll; n <= PuctConfig.batch_size), 0 when the queued work is done,
// or a negative gz_status. The next poll brings their n predictions (role r: [n][A_r], value: [n][rewards]).
extern "C" int gz_player_poll(gz_player* p, int n_pred, const float* const* h_arrays, int n_arrays, const float** h_planes) {
    if (!p) { gz_set_error("null player"); return GZ_ERR_ARG; }
    if (n_pred != p->waiting) { gz_set_error("expected %d predictions, got %d", p->waiting, n_pred); return GZ_ERR_ARG; }
    GZ_CUDA(cudaSetDevice(p->device));
    bool fresh = false;
    if (n_pred) {
        if (!h_arrays || n_arrays != p->n_roles + 1) { gz_set_error("expected %d policy arrays and one value array", p->n_roles); return GZ_ERR_ARG; }
        // role r's array is [n_pred][A_r]; the engine reads rows of concatenated roles, then the value rows
        float* hv = p->h_row + (size_t)p->rows * p->policy_stride;
        for (int i = 0; i < n_pred; i++) {
            for (int r = 0; r < p->n_roles; r++)
                memcpy(p->h_row + (size_t)i * p->policy_stride + p->policy_off[r], h_arrays[r] + (size_t)i * p->role_sizes[r], (size_t)p->role_sizes[r] * 4);
            memcpy(hv + (size_t)i * p->n_rewards, h_arrays[p->n_roles] + (size_t)i * p->n_rewards, (size_t)p->n_rewards * 4);
        }
        GZ_CUDA(cudaMemcpyAsync(p->d_policy, p->h_row, (size_t)n_pred * p->policy_stride * 4, cudaMemcpyHostToDevice, p->stream));
        GZ_CUDA(cudaMemcpyAsync(p->d_value, hv, (size_t)n_pred * p->n_rewards * 4, cudaMemcpyHostToDevice, p->stream));
        fresh = true;
    }
    p->waiting = 0;
    for (long guard = 0; guard < 100000000L; guard++) {
        if (!p->busy) {
            if (p->queue.empty()) return 0;
            int rc = player_issue_next(p);
            if (rc != GZ_OK) return rc;
        }
        int rc = player_watch_clock(p);
        if (rc != GZ_OK) return rc;
        rc = gz_puct_advance(p->eng, fresh ? p->d_policy : nullptr, fresh ? p->d_value : nullptr, fresh ? 1 : 0, 1,
                             p->d_planes, GZ_DTYPE_F32, (void*)p->stream);
        if (rc != GZ_OK) return rc;
        fresh = false;
        int st[4];
        rc = gz_puct_status(p->eng, 0, st, (void*)p->stream);
        if (rc != GZ_OK) return rc;
        if (st[2] || st[0] < 0) { gz_set_error("device player failed: error %d, cmd_done %d", st[2], st[0]); return GZ_ERR_LIMIT; }
        if (st[3]) {
            // the rows a batch hands out are the first st[3] rows of the game, in requestor order
            GZ_CUDA(cudaMemcpyAsync(p->h_planes, p->d_planes, (size_t)st[3] * p->total * 4, cudaMemcpyDeviceToHost, p->stream));
            GZ_CUDA(cudaStreamSynchronize(p->stream));
            p->waiting = st[3];
            if (h_planes) *h_planes = p->h_planes;
            return st[3];
        }
        if (st[0] == 1) {
            p->busy = false;
            if (p->current == 3) {
                p->ch_n = gz_puct_snapshot(p->eng, 0, p->max_actions, p->ch_i, p->ch_f, p->ch_moves.data(), p->ch_trav.data(),
                                           p->ch_prob.data(), p->ch_score.data(), p->ch_next.data(), (void*)p->stream);
                if (p->ch_n < 0) return p->ch_n;
                p->have_choice = true;
            }
        }
    }
    gz_set_error("device player did not finish");
    return GZ_ERR_LIMIT;
}

// The finished search's root children, as gz_puct_snapshot reports them (tests, PUCTPlayer's debugging output).
extern "C" int gz_player_last_search(gz_player* p, int* out_i16, float* out_f8, int* moves, int* trav, float* prob,
                                     float* score, float* next_prob, int cap) {
    if (!p || !p->have_choice) return 0;
    memcpy(out_i16, p->ch_i, sizeof(p->ch_i));
    memcpy(out_f8, p->ch_f, sizeof(p->ch_f));
    const int n = std::min(p->ch_n, cap);
    for (int i = 0; i < n; i++) { moves[i] = p->ch_moves[i]; trav[i] = p->ch_trav[i]; prob[i] = p->ch_prob[i]; score[i] = p->ch_score[i]; next_prob[i] = p->ch_next[i]; }
    return n;
}
