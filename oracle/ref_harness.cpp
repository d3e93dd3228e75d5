/* ORACLE harness (test infrastructure). Built by oracle/Makefile into oracle/_ref/libgzref.so
 * together with the UNMODIFIED reference sources compiled where they lie:
 *   /root/reference/src/cpp/{puct/node,puct/evaluator,scheduler,gdltransformer,selfplay,
 *                            selfplaymanager,supervisor}.cpp
 * against the shim headers in oracle/shims/ (ggplib statemachine/*.h, k273/*.h, greenlet).
 * Nothing of the reference is copied into this repository.
 *
 * This file provides what the shims only declare (greenlets on ucontext, the RNG seed counter, the
 * concrete state machine on top of oracle/propnet_ref.c) and a C driver API:
 *   - gzref_search:    one PuctEvaluator, `evals` evaluations from a given root, repeated for
 *                      `n_moves` moves with fastApplyMove in between; returns root child statistics
 *                      (the visit-count distribution parity oracle).
 *   - gzref_selfplay:  the reference Supervisor (inline or worker threads) polled in a loop, exactly
 *                      as src/ggpzero/util/cppinterface.py:78-129 does, with a synthetic network in
 *                      place of Keras; returns samples and counters (the CPU baseline, kind
 *                      "reference").
 * The synthetic network is a deterministic hash of the encoded planes (gz_synth.h), shared
 * bit-for-bit with the CUDA test kernel in csrc/gz_synth.cu.
 */
#include <sys/mman.h>
#include <ucontext.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include <greenlet/greenlet.h>
#include <k273/rng.h>
#include <statemachine/statemachine.h>

#include "gdltransformer.h"
#include "puct/config.h"
#include "puct/evaluator.h"
#include "puct/node.h"
#include "sample.h"
#include "scheduler.h"
#include "selfplay.h"
#include "selfplaymanager.h"
#include "supervisor.h"

#include "../include/gz_synth.h"

extern "C" {
#include "../include/gz_propnet_format.h"
struct pnref_sm;
pnref_sm* pnref_create(const int32_t* blob, int64_t n_words);
void pnref_destroy(pnref_sm* sm);
int pnref_n_bases(const pnref_sm* sm);
int pnref_n_roles(const pnref_sm* sm);
int pnref_role_size(const pnref_sm* sm, int r);
int pnref_state_words(const pnref_sm* sm);
void pnref_initial_state(const pnref_sm* sm, uint32_t* out);
void pnref_update_bases(pnref_sm* sm, const uint32_t* state);
int pnref_legal_count(const pnref_sm* sm, int role);
int pnref_legal(const pnref_sm* sm, int role, int i);
int pnref_is_terminal(const pnref_sm* sm);
int pnref_goal(const pnref_sm* sm, int role);
void pnref_next_state(pnref_sm* sm, const int32_t* joint_move, uint32_t* out_state);
}

// ------------------------------------------------------------------------------------------------
// RNG seed counter
static std::atomic<uint64_t> g_seed_counter{0x5EED0000ULL};
uint64_t K273::gzref_next_seed() { return g_seed_counter.fetch_add(0x9E3779B97F4A7C15ULL); }
extern "C" void gzref_seed(uint64_t s) { g_seed_counter.store(s); }

// ------------------------------------------------------------------------------------------------
// Staggered starts (measurement aid, bench.py): while enabled, every getInitialState() call — the reference makes
// one per SelfPlay slot (selfplaymanager.cpp:140) — returns a position `hash(slot) % max_ply` uniformly random
// plies into a game, so that the slots sit at all plies at once like a long-running job instead of all walking
// through the opening together. gzref_selfplay puts the true initial state back into a slot's buffer once the
// slot's first game is under way, so every later game starts normally.
static std::atomic<int> g_stagger_max_ply{0};
static std::atomic<uint64_t> g_stagger_seed{0};
static std::atomic<uint64_t> g_stagger_slot{0};
static std::atomic<uint64_t> g_stagger_total{0};     /* slots to stagger (the self-play slots of one run) */
extern "C" void gzref_set_stagger(int max_ply, uint64_t seed) {
    g_stagger_max_ply.store(max_ply);
    g_stagger_seed.store(seed);
    g_stagger_slot.store(0);
    g_stagger_total.store(0);
}
static inline uint64_t stagger_mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// ------------------------------------------------------------------------------------------------
// greenlets on ucontext
struct greenlet_t {
    ucontext_t ctx;
    std::function<void()> fn;
    greenlet_t* parent = nullptr;
    void* stack = nullptr;
    size_t stack_size = 0;
    bool dead = false;
    bool started = false;
    bool is_root = false;
};

static thread_local greenlet_t* tl_current = nullptr;
static thread_local greenlet_t* tl_root = nullptr;

static greenlet_t* root_greenlet() {
    if (!tl_root) {
        tl_root = new greenlet_t();
        tl_root->is_root = true;
        tl_root->started = true;
        tl_current = tl_root;
    }
    return tl_root;
}

greenlet_t* greenlet_current() {
    root_greenlet();
    return tl_current;
}

bool greenlet_isdead(greenlet_t* g) { return g->dead; }

static void greenlet_trampoline(unsigned lo, unsigned hi) {
    greenlet_t* g = (greenlet_t*)(((uintptr_t)hi << 32) | (uintptr_t)lo);
    g->fn();
    g->dead = true;
    greenlet_t* p = g->parent;
    while (p && p->dead) p = p->parent;
    if (!p) p = tl_root;
    tl_current = p;
    setcontext(&p->ctx);
}

greenlet_t* gzref_create_greenlet(std::function<void()> fn, greenlet_t* parent) {
    root_greenlet();
    greenlet_t* g = new greenlet_t();
    g->fn = std::move(fn);
    g->parent = parent ? parent : tl_current;
    g->stack_size = 256 * 1024;
    g->stack = mmap(nullptr, g->stack_size, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    getcontext(&g->ctx);
    g->ctx.uc_stack.ss_sp = g->stack;
    g->ctx.uc_stack.ss_size = g->stack_size;
    g->ctx.uc_link = nullptr;
    uintptr_t p = (uintptr_t)g;
    makecontext(&g->ctx, (void (*)())greenlet_trampoline, 2, (unsigned)(p & 0xffffffffu), (unsigned)(p >> 32));
    return g;
}

void greenlet_switch_to(greenlet_t* g) {
    root_greenlet();
    greenlet_t* cur = tl_current;
    if (g == cur) return;
    tl_current = g;
    g->started = true;
    swapcontext(&cur->ctx, &g->ctx);
}

// ------------------------------------------------------------------------------------------------
// concrete state machine over oracle/propnet_ref.c
struct GameTables {
    std::vector<int32_t> blob;
    int n_bases, n_roles, state_words;
    std::vector<int> role_sizes;
};

class ShimStateMachine : public GGPLib::StateMachineInterface {
public:
    explicit ShimStateMachine(const GameTables* gt) : gt(gt) {
        sm = pnref_create(gt->blob.data(), (int64_t)gt->blob.size());
        initial = newBaseState();
        pnref_initial_state(sm, (uint32_t*)initial->data);
        jm = (GGPLib::JointMove*)malloc(GGPLib::JointMove::mallocSize(gt->n_roles));
        jm->setSize(gt->n_roles);
        legal_states.resize(gt->n_roles);
        legal_lists.resize(gt->n_roles);
        for (int r = 0; r < gt->n_roles; r++) legal_lists[r].resize(gt->role_sizes[r] + 1);
        updateBases(initial);
    }
    ~ShimStateMachine() override { pnref_destroy(sm); free(initial); free(jm); }

    GGPLib::StateMachineInterface* dupe() const override { return new ShimStateMachine(gt); }
    GGPLib::BaseState* newBaseState() const override {
        GGPLib::BaseState* bs = (GGPLib::BaseState*)malloc(GGPLib::BaseState::mallocSize(gt->n_bases));
        bs->init(gt->n_bases);
        return bs;
    }
    const GGPLib::BaseState* getInitialState() const override {
        const int max_ply = g_stagger_max_ply.load();
        if (max_ply <= 0) return initial;
        return const_cast<ShimStateMachine*>(this)->staggeredState(max_ply);
    }
    const GGPLib::BaseState* trueInitialState() const { return initial; }
    /* a fresh buffer (owned by the caller's slot for the life of the process) holding a random-playout position */
    GGPLib::BaseState* staggeredState(int max_ply) {
        const uint64_t slot = g_stagger_slot.fetch_add(1);
        if (slot >= g_stagger_total.load()) return const_cast<GGPLib::BaseState*>(initial);
        uint64_t rs = stagger_mix(g_stagger_seed.load() ^ (slot * 0xD1B54A32D192ED03ULL));
        const int plies = (int)(stagger_mix(rs) % (uint64_t)max_ply);
        GGPLib::BaseState* cur = newBaseState();
        GGPLib::BaseState* nxt = newBaseState();
        cur->assign(initial);
        for (int p = 0; p < plies; p++) {
            updateBases(cur);
            for (int r = 0; r < gt->n_roles; r++) {
                rs = stagger_mix(rs);
                jm->set(r, legal_states[r].legals[rs % (uint64_t)legal_states[r].count]);
            }
            nextState(jm, nxt);
            updateBases(nxt);
            if (isTerminal()) { cur->assign(initial); break; }   // the random game ended early: start normally
            std::swap(cur, nxt);
        }
        free(nxt);
        updateBases(initial);
        return cur;
    }
    void updateBases(const GGPLib::BaseState* bs) override {
        pnref_update_bases(sm, (const uint32_t*)bs->data);
        for (int r = 0; r < gt->n_roles; r++) {
            const int c = pnref_legal_count(sm, r);
            for (int i = 0; i < c; i++) legal_lists[r][i] = pnref_legal(sm, r, i);
            legal_states[r].count = c;
            legal_states[r].legals = legal_lists[r].data();
        }
    }
    void reset() override { updateBases(initial); }
    GGPLib::LegalState* getLegalState(int ri) override { return &legal_states[ri]; }
    GGPLib::JointMove* getJointMove() override { return jm; }
    void nextState(const GGPLib::JointMove* move, GGPLib::BaseState* out) override {
        int32_t m[GZ_PN_MAX_ROLES];
        for (int r = 0; r < gt->n_roles; r++) m[r] = move->get(r);
        std::vector<uint32_t> tmp(gt->state_words);
        pnref_next_state(sm, m, tmp.data());
        std::memset(out->data, 0, out->byte_count);
        std::memcpy(out->data, tmp.data(), sizeof(uint32_t) * gt->state_words);
    }
    bool isTerminal() const override { return pnref_is_terminal(sm) != 0; }
    int getGoalValue(int ri) override { return pnref_goal(sm, ri); }
    int getRoleCount() const override { return gt->n_roles; }
    const char* legalToMove(int, int) const override { return "?"; }

private:
    const GameTables* gt;
    pnref_sm* sm;
    GGPLib::BaseState* initial;
    GGPLib::JointMove* jm;
    std::vector<GGPLib::LegalState> legal_states;
    std::vector<std::vector<int>> legal_lists;
};

// ------------------------------------------------------------------------------------------------
using namespace GGPZero;

struct TransformerSpec {
    int channel_size, channels_per_state, num_control_channels, num_prev_states, num_rewards;
    std::vector<int> policy_sizes;
    std::vector<int> board_base, board_off, ctrl_base, ctrl_channel;
    std::vector<float> ctrl_value;
    GdlBasesTransformer* make() const {
        std::vector<int> ps = policy_sizes;
        GdlBasesTransformer* t = new GdlBasesTransformer(channel_size, channels_per_state, num_control_channels,
                                                         num_prev_states, num_rewards, ps);
        for (size_t i = 0; i < board_base.size(); i++) t->addBoardBase(board_base[i], board_off[i]);
        for (size_t i = 0; i < ctrl_base.size(); i++) t->addControlBase(ctrl_base[i], ctrl_channel[i], ctrl_value[i]);
        return t;
    }
};

struct GzRef {
    GameTables gt;
    TransformerSpec ts;
};

extern "C" void* gzref_create(const int32_t* blob, int64_t n_words, int channel_size, int channels_per_state,
                              int num_control_channels, int num_prev_states, int num_rewards,
                              const int* policy_sizes, int n_policies) {
    GzRef* g = new GzRef();
    g->gt.blob.assign(blob, blob + n_words);
    pnref_sm* sm = pnref_create(blob, n_words);
    if (!sm) { delete g; return nullptr; }
    g->gt.n_bases = pnref_n_bases(sm);
    g->gt.n_roles = pnref_n_roles(sm);
    g->gt.state_words = pnref_state_words(sm);
    for (int r = 0; r < g->gt.n_roles; r++) g->gt.role_sizes.push_back(pnref_role_size(sm, r));
    pnref_destroy(sm);
    g->ts.channel_size = channel_size;
    g->ts.channels_per_state = channels_per_state;
    g->ts.num_control_channels = num_control_channels;
    g->ts.num_prev_states = num_prev_states;
    g->ts.num_rewards = num_rewards;
    g->ts.policy_sizes.assign(policy_sizes, policy_sizes + n_policies);
    return g;
}
extern "C" void gzref_destroy(void* h) { delete (GzRef*)h; }
extern "C" void gzref_add_board_base(void* h, int b, int off) {
    GzRef* g = (GzRef*)h; g->ts.board_base.push_back(b); g->ts.board_off.push_back(off);
}
extern "C" void gzref_add_control_base(void* h, int b, int c, float v) {
    GzRef* g = (GzRef*)h; g->ts.ctrl_base.push_back(b); g->ts.ctrl_channel.push_back(c); g->ts.ctrl_value.push_back(v);
}

/* The reference's own encoder: GdlBasesTransformer::toChannels (gdltransformer.cpp:22-52), compiled unmodified, on
 * `n` packed states. prev[(i * max_prev + k) * SW ..] = k-th previous state of row i (nearest first), of which the
 * first prev_count[i] (all max_prev when prev_count is null) are handed over, as PuctNodeRequest::add does while
 * walking the parents (node.cpp:449-461). out: [n][totalSize()] float32. */
extern "C" int gzref_encode(void* h, int64_t n, const uint32_t* states, const uint32_t* prev, const uint8_t* prev_count,
                            int max_prev, float* out) {
    GzRef* g = (GzRef*)h;
    ShimStateMachine sm(&g->gt);
    const GdlBasesTransformer* tr = g->ts.make();
    const int SW = g->gt.state_words, total = tr->totalSize();
    GGPLib::BaseState* cur = sm.newBaseState();
    std::vector<GGPLib::BaseState*> pool;
    for (int k = 0; k < max_prev; k++) pool.push_back(sm.newBaseState());
    for (int64_t i = 0; i < n; i++) {
        std::memcpy(cur->data, states + i * SW, sizeof(uint32_t) * SW);
        std::vector<const GGPLib::BaseState*> ps;
        const int cnt = prev ? (prev_count ? prev_count[i] : max_prev) : 0;
        for (int k = 0; k < cnt && k < max_prev; k++) {
            std::memcpy(pool[k]->data, prev + (i * max_prev + k) * SW, sizeof(uint32_t) * SW);
            ps.push_back(pool[k]);
        }
        tr->toChannels(cur, ps, out + i * total);
    }
    free(cur);
    for (auto* p : pool) free(p);
    delete tr;
    return total;
}

/* PuctConfig from a flat array, field order = GZ_PUCT_CONF_* below (mirrors config.h:11-53). */
enum {
    PC_PUCT_CONSTANT = 0, PC_PUCT_CONSTANT_ROOT, PC_DIRICHLET_NOISE_PCT, PC_NOISE_POLICY_SQUASH_PCT,
    PC_NOISE_POLICY_SQUASH_PROB, PC_CHOOSE, PC_RANDOM_SCALE, PC_TEMPERATURE, PC_DEPTH_TEMPERATURE_START,
    PC_DEPTH_TEMPERATURE_INCREMENT, PC_DEPTH_TEMPERATURE_STOP, PC_DEPTH_TEMPERATURE_MAX,
    PC_FPU_PRIOR_DISCOUNT, PC_FPU_PRIOR_DISCOUNT_ROOT, PC_TOP_VISITS_BEST_GUESS_CONVERGE_RATIO,
    PC_THINK_TIME, PC_CONVERGED_VISITS, PC_BATCH_SIZE, PC_USE_LEGALS_COUNT_DRAW, PC_BACKUP_FINALISED,
    PC_LOOKUP_TRANSPOSITIONS, PC_EVALUATION_MULTIPLIER_TO_CONVERGENCE, PC_COUNT
};

static PuctConfig* make_puct_config(const double* c) {
    PuctConfig* p = new PuctConfig();
    p->verbose = false;
    p->puct_constant = (float)c[PC_PUCT_CONSTANT];
    p->puct_constant_root = (float)c[PC_PUCT_CONSTANT_ROOT];
    p->dirichlet_noise_pct = (float)c[PC_DIRICHLET_NOISE_PCT];
    p->noise_policy_squash_pct = (float)c[PC_NOISE_POLICY_SQUASH_PCT];
    p->noise_policy_squash_prob = (float)c[PC_NOISE_POLICY_SQUASH_PROB];
    p->choose = c[PC_CHOOSE] > 0.5 ? ChooseFn::choose_temperature : ChooseFn::choose_top_visits;
    p->max_dump_depth = 0;
    p->random_scale = (float)c[PC_RANDOM_SCALE];
    p->temperature = (float)c[PC_TEMPERATURE];
    p->depth_temperature_start = (int)c[PC_DEPTH_TEMPERATURE_START];
    p->depth_temperature_increment = (float)c[PC_DEPTH_TEMPERATURE_INCREMENT];
    p->depth_temperature_stop = (int)c[PC_DEPTH_TEMPERATURE_STOP];
    p->depth_temperature_max = (float)c[PC_DEPTH_TEMPERATURE_MAX];
    p->fpu_prior_discount = (float)c[PC_FPU_PRIOR_DISCOUNT];
    p->fpu_prior_discount_root = (float)c[PC_FPU_PRIOR_DISCOUNT_ROOT];
    p->top_visits_best_guess_converge_ratio = (float)c[PC_TOP_VISITS_BEST_GUESS_CONVERGE_RATIO];
    p->think_time = (float)c[PC_THINK_TIME];
    p->converged_visits = (int)c[PC_CONVERGED_VISITS];
    p->batch_size = (int)c[PC_BATCH_SIZE];
    p->use_legals_count_draw = (int)c[PC_USE_LEGALS_COUNT_DRAW];
    p->backup_finalised = c[PC_BACKUP_FINALISED] > 0.5;
    p->lookup_transpositions = c[PC_LOOKUP_TRANSPOSITIONS] > 0.5;
    p->evaluation_multiplier_to_convergence = (float)c[PC_EVALUATION_MULTIPLIER_TO_CONVERGENCE];
    return p;
}

/* Runs the synthetic network on `n` encoded rows and fills a PredictDoneEvent's buffers. */
static void synth_predict(const TransformerSpec& ts, uint64_t nn_seed, const float* planes, int n,
                          std::vector<std::vector<float>>& policies, std::vector<float>& values) {
    const int total = ts.channel_size * (ts.channels_per_state * (ts.num_prev_states + 1) + ts.num_control_channels);
    for (size_t r = 0; r < ts.policy_sizes.size(); r++) policies[r].resize((size_t)n * ts.policy_sizes[r]);
    values.resize((size_t)n * ts.num_rewards);
    for (int i = 0; i < n; i++) {
        const uint64_t h = gz_synth_hash_planes_f32(planes + (size_t)i * total, total, nn_seed);
        for (size_t r = 0; r < ts.policy_sizes.size(); r++)
            for (int a = 0; a < ts.policy_sizes[r]; a++)
                policies[r][(size_t)i * ts.policy_sizes[r] + a] = gz_synth_policy(h, (uint32_t)r, (uint32_t)a);
        gz_synth_values(h, ts.num_rewards, &values[(size_t)i * ts.num_rewards]);
    }
}

/* One evaluator, sequential search; for each of n_moves: `evals` evaluations, record the root's
 * children, then apply the chosen move. Outputs (per move m, child c < max_children):
 *   out_counts[m] = num_children, out_move[m*max_children+c] = lead-role action index,
 *   out_trav = traversals, out_prob = policy_prob, out_score = child node current score for the
 *   lead role (-1 when unexpanded), out_choice[m] = index of the chosen child,
 *   out_root[m*4+{0,1,2,3}] = root visits, lead role, root score(lead), #nodes.
 * Returns the number of moves actually searched (stops at terminal). */
/* scores of the chosen child's node for every role, per move of the last gzref_search (what
 * Player::puctPlayerGetMove reports for the role it is asked about, player.cpp:105-119) */
static std::vector<float> g_last_choice_scores;
extern "C" int gzref_last_choice_scores(float* out, int cap) {
    const int n = (int)g_last_choice_scores.size() < cap ? (int)g_last_choice_scores.size() : cap;
    for (int i = 0; i < n; i++) out[i] = g_last_choice_scores[i];
    return n;
}

/* PuctNode::debug of the root's first three children in debug order, per move of the last gzref_search
 * (Player::treeDebugInfo, player.cpp:137-153): per line 24 floats = valid, score, lead, move, n pairs, 10 x (lead, move) */
static std::vector<float> g_last_debug;
/* Player::balanceNode (player.cpp:121-135): when > 0, every search of gzref_search is followed by
 * balanceFirstMoves(max_moves) (evaluator.cpp:282-339) and the root's traversals after it are kept */
static int g_balance_moves = 0;
static std::vector<int> g_last_balance;          /* per move: num_children, then that many traversal counts */
extern "C" void gzref_set_balance(int max_moves) { g_balance_moves = max_moves; }
extern "C" int gzref_last_balance(int* out, int cap) {
    const int n = (int)g_last_balance.size() < cap ? (int)g_last_balance.size() : cap;
    for (int i = 0; i < n; i++) out[i] = g_last_balance[i];
    return n;
}
extern "C" int gzref_last_debug(float* out, int cap) {
    const int n = (int)g_last_debug.size() < cap ? (int)g_last_debug.size() : cap;
    for (int i = 0; i < n; i++) out[i] = g_last_debug[i];
    return n;
}

extern "C" int gzref_search(void* h, const double* puct_conf, const uint32_t* root_state, int evals, int n_moves,
                            uint64_t nn_seed, int max_children, int* out_counts, int* out_move, int* out_trav,
                            float* out_prob, float* out_score, int* out_choice, float* out_root) {
    GzRef* g = (GzRef*)h;
    ShimStateMachine* sm = new ShimStateMachine(&g->gt);
    const GdlBasesTransformer* tr = g->ts.make();
    PuctConfig* conf = make_puct_config(puct_conf);
    /* Player (player.cpp:27-36): the scheduler's batch is PuctConfig.batch_size; 1 = the self-play setting */
    NetworkScheduler* sched = new NetworkScheduler(tr, conf->batch_size > 1 ? conf->batch_size : 1);
    sched->createMainLoop();
    PuctEvaluator* pe = new PuctEvaluator(sm, sched, tr);
    pe->updateConf(conf);
    GGPLib::BaseState* root_bs = sm->newBaseState();
    if (root_state) std::memcpy(root_bs->data, root_state, sizeof(uint32_t) * g->gt.state_words);
    else root_bs->assign(sm->getInitialState());
    int done_moves = 0;
    const int R = g->gt.n_roles;
    g_last_choice_scores.clear();
    g_last_debug.clear();
    g_last_balance.clear();
    auto f = [&]() {
        pe->reset(0);
        PuctNode* node = pe->establishRoot(root_bs);
        for (int m = 0; m < n_moves; m++) {
            if (node->isTerminal()) break;
            const PuctNodeChild* choice = pe->onNextMove(evals);
            const int lead = node->lead_role_index >= 0 ? node->lead_role_index : 0;
            out_counts[m] = node->num_children;
            out_choice[m] = -1;
            for (int c = 0; c < node->num_children && c < max_children; c++) {
                const PuctNodeChild* ch = node->getNodeChild(R, c);
                out_move[m * max_children + c] = ch->move.get(lead);
                out_trav[m * max_children + c] = (int)ch->traversals;
                out_prob[m * max_children + c] = ch->policy_prob;
                out_score[m * max_children + c] = ch->to_node ? ch->to_node->getCurrentScore(lead) : -1.0f;
                if (ch == choice) out_choice[m] = c;
            }
            out_root[m * 4 + 0] = (float)node->visits;
            out_root[m * 4 + 1] = (float)node->lead_role_index;
            out_root[m * 4 + 2] = node->getCurrentScore(lead);
            out_root[m * 4 + 3] = (float)pe->nodeCount();
            done_moves = m + 1;
            for (int r = 0; r < 4; r++)
                g_last_choice_scores.push_back(r < R && choice->to_node ? choice->to_node->getCurrentScore(r) : -1.0f);
            for (int ii = 0; ii < 3; ii++) {
                PuctNodeDebug info;
                info.score = -2.0f; info.lead_role_index = -2; info.move_index = -2;
                if (ii < node->num_children) PuctNode::debug(node, ii, 10, info);
                float line[25] = {0};
                line[0] = info.lead_role_index != -2 ? 1.0f : 0.0f;
                line[1] = info.score; line[2] = (float)info.lead_role_index; line[3] = (float)info.move_index;
                line[4] = (float)info.variation.size();
                for (size_t v = 0; v < info.variation.size() && v < 10; v++) {
                    line[5 + 2 * v] = (float)info.variation[v].lead_role_index;
                    line[6 + 2 * v] = (float)info.variation[v].move_index;
                }
                for (int q = 0; q < 25; q++) g_last_debug.push_back(line[q]);
            }
            if (g_balance_moves > 0) {
                pe->balanceFirstMoves(g_balance_moves);
                g_last_balance.push_back(node->num_children);
                for (int c = 0; c < node->num_children; c++) g_last_balance.push_back((int)node->getNodeChild(R, c)->traversals);
            }
            node = pe->fastApplyMove(choice);
            if (getenv("GZREF_TRACE")) fprintf(stderr, "A %d nodes %d root_visits %d\n", m, (int)pe->nodeCount(), (int)node->visits);
        }
    };
    sched->addRunnable(f);
    PredictDoneEvent pred;
    ReadyEvent ready;
    std::vector<std::vector<float>> policies(g->ts.policy_sizes.size());
    std::vector<float> values;
    pred.pred_count = 0;
    pred.final_scores = nullptr;
    const int total = tr->totalSize();
    while (true) {
        sched->poll(&pred, &ready);
        if (ready.buf_count == 0) break;
        const int n = ready.buf_count / total;
        if (getenv("GZREF_TRACE")) fprintf(stderr, "B %d\n", n);
        synth_predict(g->ts, nn_seed, ready.channel_buf, n, policies, values);
        pred.policies.resize(policies.size());
        for (size_t r = 0; r < policies.size(); r++) pred.policies[r] = policies[r].data();
        pred.final_scores = values.data();
        pred.pred_count = n;
    }
    free(root_bs);
    /* evaluator / scheduler / transformer are intentionally leaked: NetworkScheduler deletes the
       transformer it was given (scheduler.cpp:65-68) and the evaluator's destructor expects a live
       scheduler; this is a short-lived test process. */
    return done_moves;
}

/* Plies played so far by every SelfPlay of a manager. The build compiles this translation unit and
 * the reference sources with -Dprivate=public (no source is modified) so that the harness can read
 * SelfPlay::match_count and PuctEvaluator::game_depth. Reads race benignly with the worker threads. */
struct MoveCounter {
    std::vector<int> last_match, last_depth;
    std::vector<char> restored;
    const GGPLib::BaseState* true_initial = nullptr;   // set when the slots were given staggered starts
    double finished_plies = 0;
    double count(const std::vector<SelfPlayManager*>& mans) {
        size_t idx = 0;
        double cur = 0;
        for (SelfPlayManager* m : mans) {
            for (SelfPlay* sp : m->self_plays) {
                if (idx >= last_match.size()) { last_match.push_back(sp->match_count); last_depth.push_back(0); restored.push_back(0); }
                const int mc = sp->match_count;
                const int d = sp->pe->game_depth;
                if (true_initial && !restored[idx] && mc >= 1 && d >= 1) {
                    /* the slot's first game has left its (staggered) root behind: later games start normally */
                    const_cast<GGPLib::BaseState*>(sp->initial_state)->assign(true_initial);
                    restored[idx] = 1;
                }
                if (mc != last_match[idx]) {
                    finished_plies += last_depth[idx];
                    last_match[idx] = mc;
                    last_depth[idx] = 0;
                }
                if (mc == last_match[idx] && d >= last_depth[idx]) last_depth[idx] = d;
                cur += last_depth[idx];
                idx++;
            }
        }
        return finished_plies + cur;
    }
};

/* SelfPlayConfig flat layout */
enum {
    SC_OSCILLATE_SAMPLING_PCT = 0, SC_TEMPERATURE_FOR_POLICY, SC_EVALS_PER_MOVE, SC_RESIGN0_SCORE_PROBABILITY,
    SC_RESIGN0_PCT, SC_RESIGN1_SCORE_PROBABILITY, SC_RESIGN1_PCT, SC_RUN_TO_END_PCT, SC_RUN_TO_END_EVALS,
    SC_RUN_TO_END_EARLY_SCORE, SC_RUN_TO_END_MINIMUM_GAME_DEPTH, SC_ABORT_MAX_LENGTH, SC_COUNT
};

/* Reference self-play: Supervisor with `num_workers` worker threads (<=0: inline), `batch_size`
 * games per manager, polled until `target_samples` samples were fetched or `max_seconds` elapsed.
 * out_stats: [0] samples, [1] evaluations (NN rows), [2] polls, [3] seconds,
 *            [4] sum of sample depth (for sanity), [5] sampled moves == samples.
 * Moves/s is derived by the caller from game_length of completed games:
 *            [6] completed games (distinct match ids seen), [7] total plies of completed games,
 *            [8] in: warm-up seconds / out: plies played in total, [9] window seconds,
 *            [10] plies in the window, [11] evaluations in the window. */
extern "C" int gzref_selfplay(void* h, const double* sp_conf, const double* puct_conf,
                              const double* run_to_end_puct_conf, int batch_size, int num_workers,
                              uint64_t nn_seed, int64_t target_samples, double max_seconds, double* out_stats,
                              int32_t* out_samples, int64_t out_cap, int srec) {
    GzRef* g = (GzRef*)h;
    ShimStateMachine* sm = new ShimStateMachine(&g->gt);
    const GdlBasesTransformer* tr = g->ts.make();
    SelfPlayConfig* sc = new SelfPlayConfig();
    sc->oscillate_sampling_pct = (float)sp_conf[SC_OSCILLATE_SAMPLING_PCT];
    sc->temperature_for_policy = (float)sp_conf[SC_TEMPERATURE_FOR_POLICY];
    sc->puct_config = make_puct_config(puct_conf);
    sc->evals_per_move = (int)sp_conf[SC_EVALS_PER_MOVE];
    sc->resign0_score_probability = (float)sp_conf[SC_RESIGN0_SCORE_PROBABILITY];
    sc->resign0_pct = (float)sp_conf[SC_RESIGN0_PCT];
    sc->resign1_score_probability = (float)sp_conf[SC_RESIGN1_SCORE_PROBABILITY];
    sc->resign1_pct = (float)sp_conf[SC_RESIGN1_PCT];
    sc->abort_max_length = (int)sp_conf[SC_ABORT_MAX_LENGTH];
    sc->number_repeat_states_draw = -1;
    sc->repeat_states_score = 0.5f;
    sc->run_to_end_pct = (float)sp_conf[SC_RUN_TO_END_PCT];
    sc->run_to_end_evals = (int)sp_conf[SC_RUN_TO_END_EVALS];
    sc->run_to_end_puct_config = make_puct_config(run_to_end_puct_conf);
    sc->run_to_end_early_score = (float)sp_conf[SC_RUN_TO_END_EARLY_SCORE];
    sc->run_to_end_minimum_game_depth = (int)sp_conf[SC_RUN_TO_END_MINIMUM_GAME_DEPTH];

    /* worker threads create their SelfPlay slots asynchronously (supervisor.cpp:196-207): staggering stays armed
       for exactly the slots of this run */
    g_stagger_total.store((uint64_t)batch_size * (uint64_t)(num_workers <= 0 ? 1 : 2 * num_workers));
    Supervisor* sup = new Supervisor(sm, tr, batch_size, "gzref");
    if (num_workers <= 0) {
        sup->createInline(sc);
    } else {
        for (int i = 0; i < num_workers; i++) sup->createWorkers(sc);
    }
    std::vector<std::vector<float>> policies(g->ts.policy_sizes.size());
    std::vector<float> values;
    std::vector<float*> data;
    const int total = tr->totalSize();
    int n_pred = 0;
    int64_t n_samples = 0, n_evals = 0, n_polls = 0;
    double depth_sum = 0, plies_sum = 0;
    int64_t games = 0;
    std::string last_id;
    std::vector<SelfPlayManager*> managers;
    if (sup->inline_sp_manager) managers.push_back(sup->inline_sp_manager);
    for (SelfPlayWorker* w : sup->self_play_workers) { managers.push_back(w->man0); managers.push_back(w->man1); }
    MoveCounter mc;
    if (g_stagger_max_ply.load() > 0) mc.true_initial = sm->trueInitialState();
    double moves_now = 0, moves_at_warm = 0, evals_at_warm = 0, t_warm = -1;
    const double warm_seconds = out_stats[8];     /* in: seconds to exclude from the rates */
    const double t0 = K273::get_time();
    while (true) {
        const ReadyEvent* ev = sup->poll(n_pred, data);
        n_polls++;
        if (ev->buf_count == 0) break;
        const int n = ev->buf_count / total;
        n_evals += n;
        synth_predict(g->ts, nn_seed, ev->channel_buf, n, policies, values);
        data.clear();
        for (size_t r = 0; r < policies.size(); r++) data.push_back(policies[r].data());
        data.push_back(values.data());
        n_pred = n;
        if ((n_polls & 15) == 0) {
            moves_now = mc.count(managers);
            if (t_warm < 0 && K273::get_time() - t0 >= warm_seconds) {
                t_warm = K273::get_time();
                moves_at_warm = moves_now;
                evals_at_warm = (double)n_evals;
            }
        }
        if ((n_polls & 63) == 0) {
            std::vector<Sample*> ss = sup->getSamples();
            for (Sample* s : ss) {
                if (out_samples && n_samples < out_cap) {
                    /* same record layout as csrc/gz_puct_core.h (SR_*) */
                    int32_t* rec = out_samples + n_samples * srec;
                    float* fr = (float*)rec;
                    const int SW = g->gt.state_words;
                    int slot = 0, match = 0;
                    {
                        const std::string& id = s->match_identifier;
                        size_t p2 = id.rfind('_');
                        size_t p1 = id.rfind('_', p2 - 1);
                        match = atoi(id.c_str() + p2 + 1);
                        slot = atoi(id.c_str() + p1 + 1);
                    }
                    rec[0] = 1; rec[1] = s->depth; rec[2] = s->game_length; rec[3] = slot; rec[4] = match;
                    rec[5] = s->has_resigned; rec[6] = s->resign_false_positive; rec[7] = s->starting_sample_depth;
                    rec[8] = s->resultant_puct_visits; rec[9] = s->lead_role_index;
                    rec[10] = (int)s->prev_states.size(); rec[11] = (int)s->policies[s->lead_role_index].size();
                    for (int r = 0; r < 4; r++) {
                        fr[12 + r] = r < (int)s->final_score.size() ? s->final_score[r] : 0.0f;
                        fr[16 + r] = r < (int)s->resultant_puct_score.size() ? s->resultant_puct_score[r] : 0.0f;
                        rec[20 + r] = r < (int)s->policies.size() && !s->policies[r].empty() ? s->policies[r][0].first : 0;
                    }
                    memcpy(rec + 24, s->state->data, 4 * SW);
                    if (!s->prev_states.empty()) memcpy(rec + 24 + SW, s->prev_states[0]->data, 4 * SW);
                    else memset(rec + 24 + SW, 0, 4 * SW);
                    const Sample::Policy& pol = s->policies[s->lead_role_index];
                    for (size_t i = 0; i < pol.size(); i++) {
                        rec[24 + 2 * SW + 2 * i] = pol[i].first;
                        fr[24 + 2 * SW + 2 * i + 1] = pol[i].second;
                    }
                }
                n_samples++;
                depth_sum += s->depth;
                if (s->match_identifier != last_id) {   // samples of a game arrive contiguously
                    last_id = s->match_identifier;
                    games++;
                    plies_sum += s->game_length;
                }
                free(s->state);
                for (auto* p : s->prev_states) free(p);
                delete s;
            }
            if ((target_samples > 0 && n_samples >= target_samples) || K273::get_time() - t0 > max_seconds) break;
        }
    }
    out_stats[0] = (double)n_samples;
    out_stats[1] = (double)n_evals;
    out_stats[2] = (double)n_polls;
    out_stats[3] = K273::get_time() - t0;
    out_stats[4] = depth_sum;
    out_stats[5] = (double)n_samples;
    out_stats[6] = (double)games;
    out_stats[7] = plies_sum;
    moves_now = mc.count(managers);
    out_stats[8] = moves_now;                                   /* plies played in total */
    out_stats[9] = t_warm < 0 ? 0 : K273::get_time() - t_warm;  /* measured window (after warm-up) */
    out_stats[10] = moves_now - moves_at_warm;                  /* plies in the window */
    out_stats[11] = (double)n_evals - evals_at_warm;            /* evaluations in the window */
    /* Supervisor::~Supervisor does not stop its workers ("for now kill -9", supervisor.cpp:34-40):
       worker threads keep running detached; callers run this in a short-lived process. */
    return 0;
}
