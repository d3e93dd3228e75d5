// gz_puct.cu — batched PUCT self-play engine on the device.
//
// G independent game trees live in SoA node / child arenas in HBM (gz_puct_core.h; re-rooting
// compacts a game's arena in place). One engine step ("advance") is a fixed sequence of launches on one stream:
//
//   puct_pre_kernel    one warp per game: consume the network's answer (reply + backup), run the
//                      self-play controller and select down the tree until the game needs a new node
//   puct_reroot_kernel one CTA per game that just played a move: in-place compaction of the kept subtree
//   pn_kernel          (gz_propnet.cu) next state, legal masks, terminal, goals for every request
//   puct_post_kernel   one warp per game: create the node (children from the legal mask)
//   [repeat the three launches `sub_iters` times for games that did not need the network]
//   encode_kernel      (gz_encoder.cu) bf16 NCHW planes for every game waiting for the network
//
// then the caller runs the policy/value network on the planes (a PyTorch module on the same stream;
// nothing leaves the device) and hands the outputs to the next advance. The sequence has fixed
// shapes and no host synchronisation, so it can be captured in a CUDA graph.
//
// Replaces: NetworkScheduler's greenlet batching (src/cpp/scheduler.cpp), SelfPlayManager /
// SelfPlay / Supervisor (src/cpp/selfplaymanager.cpp, selfplay.cpp, supervisor.cpp) and the
// PuctEvaluator they drive (src/cpp/puct/evaluator.cpp); the search semantics are restated in
// gz_puct_core.h and are bit-exact with the reference's own C++ (tests/test_puct_hostsim.py on the
// CPU build of the same source, tests/test_gpu_puct.py on the device).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/gz_b200.h"
#include "../../include/gz_synth.h"
#include "gz_common.cuh"
#include "gz_puct_layout.h"
#include "gz_transformer.h"

extern "C" const uint32_t* gz_propnet_init_words(const gz_propnet* pn);   // gz_propnet.cu (device pointer)
extern "C" int64_t gz_propnet_wide_scratch_bytes(const gz_propnet* pn);
extern "C" int gz_propnet_goal_chunks(const gz_propnet* pn);
extern "C" int gz_sm_eval_devcount(gz_propnet* pn, int64_t n_max, const int* d_n, const uint32_t* d_states,
                                   uint32_t* d_legal, uint8_t* d_terminal, uint8_t* d_goals, int goals_mode,
                                   void* stream, uint32_t* d_scratch);
#define GZ_GOALS_NEVER_INTERNAL 2

extern "C" int gz_sm_transition_scratch(gz_propnet* pn, int64_t n, const uint32_t* d_states, const int16_t* d_moves,
                                        uint32_t* d_next, uint32_t* d_legal, uint8_t* d_terminal, uint8_t* d_goals,
                                        int goals_mode, void* stream, uint32_t* d_scratch);
extern "C" int gz_encode_layout(gz_transformer* t, int64_t n, int state_words, const uint32_t* d_states,
                                const uint32_t* d_prev, const uint8_t* d_prev_count, int max_prev, void* d_out,
                                int out_dtype, int c_pad, void* stream);

#define PUCT_WARPS_PER_BLOCK 4

// All 32 lanes of a warp run the controller of their game redundantly (identical loads, identical
// stores), which keeps the warp converged and lets the lanes cooperate where it pays: staging a
// node's children and their statistics in shared memory (gzp_cache_fill), gathering the policy row
// (gzp_reply), copying sample records. Atomics are issued by lane 0 only.
#define PUCT_CTL_WORDS ((int)((sizeof(GameCtl) + 3) / 4))
__global__ void __launch_bounds__(PUCT_WARPS_PER_BLOCK * 32, 8)
puct_pre_kernel(const __grid_constant__ PuctEngine E, int nn_fresh, int cache_words, int* __restrict__ rr_list,
                int* __restrict__ rr_count) {
    extern __shared__ __align__(16) uint32_t puct_smem[];
    // per warp: the game's controller record and playout path live in shared memory while the kernel
    // runs (they are read and written dozens of times per playout; through global memory every
    // read-after-write is an L2 round trip), followed by the child cache
    __shared__ __align__(16) uint32_t ctl_sm[PUCT_WARPS_PER_BLOCK][PUCT_CTL_WORDS];
    __shared__ int32_t ord_sm[PUCT_WARPS_PER_BLOCK][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * PUCT_WARPS_PER_BLOCK + warp;
    if (g >= E.G) return;
    uint32_t* cmem = cache_words > 0 ? puct_smem + (size_t)warp * (cache_words + 2 * E.MAXD) : nullptr;
    GameView V = gzp_view(&E, g, lane, cmem);
    uint32_t* gctl = reinterpret_cast<uint32_t*>(&E.ctl[g]);
    for (int i = lane; i < PUCT_CTL_WORDS; i += 32) ctl_sm[warp][i] = gctl[i];
    __syncwarp();
    V.c = reinterpret_cast<GameCtl*>(ctl_sm[warp]);
    V.ord = ord_sm[warp];
    int32_t *gpn = V.pn, *gpc = V.pc;
    if (cmem && E.EB == 1) {    // (intra-tree workers: every coroutine has its own path, in global memory)
        int32_t* spn = reinterpret_cast<int32_t*>(cmem + cache_words);
        int32_t* spc = spn + E.MAXD;
        const int pl = V.c->path_len;
        for (int i = lane; i < pl; i += 32) { spn[i] = gpn[i]; spc[i] = gpc[i]; }
        __syncwarp();
        V.pn = spn; V.pc = spc;
    }
    const long long dbg_t0 = E.dbg_cycles ? clock64() : 0;
    const int dbg_phase0 = V.c->phase;
    const unsigned long long dbg_po0 = V.c->total_playouts, dbg_mv0 = V.c->total_moves;
    gzp_step_pre(V, nn_fresh != 0);
    __syncwarp();
    if (E.dbg_cycles && lane == 0) {
        long long* d = E.dbg_cycles + (int64_t)g * 4;
        d[0] = clock64() - dbg_t0;
        d[1] = dbg_phase0 * 256 + V.c->phase;
        d[2] = (long long)(V.c->total_playouts - dbg_po0);
        d[3] = (long long)(V.c->total_moves - dbg_mv0) * 65536 + V.c->path_len * 256 + (V.c->req & 255);
    }
    // games that just played a move ask for their tree to be re-rooted: hand their ids to the re-root kernel
    if (lane == 0 && V.c->req == REQ_REROOT) rr_list[atomicAdd(rr_count, 1)] = g;
    if (cmem && E.EB == 1) {
        const int pl = V.c->path_len;
        for (int i = lane; i < pl; i += 32) { gpn[i] = V.pn[i]; gpc[i] = V.pc[i]; }
    }
    for (int i = lane; i < PUCT_CTL_WORDS; i += 32) gctl[i] = ctl_sm[warp][i];
}

__global__ void __launch_bounds__(PUCT_WARPS_PER_BLOCK * 32)
puct_post_kernel(const __grid_constant__ PuctEngine E, int which /* 0 all, 1 non-terminal states, 2 terminal states */) {
    const int g = blockIdx.x * PUCT_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (g >= E.G || (threadIdx.x & 31) != 0) return;
    if (which != 0 && (E.x_term[g] != 0) != (which == 2)) return;
    GameView V = gzp_view(&E, g);
    gzp_step_post(V);
}


// Re-rooting: one CTA per game that asked for it (a handful per step); in-place parallel
// mark-compact of the kept subtree (gzp_reroot_compact).
#define REROOT_THREADS 1024
struct GzpCta {
    int* sm;   // [2 * (REROOT_THREADS / 32)]
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ void scan2(int a, int b, int& ea, int& eb, int& ta, int& tb) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        int ia = a, ib = b;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int xa = __shfl_up_sync(0xffffffffu, ia, d), xb = __shfl_up_sync(0xffffffffu, ib, d);
            if (lane >= d) { ia += xa; ib += xb; }
        }
        if (lane == 31) { sm[warp] = ia; sm[nw + warp] = ib; }
        __syncthreads();
        int pa = 0, pb = 0, sa = 0, sb = 0;
        for (int w = 0; w < nw; w++) {
            const int va = sm[w], vb = sm[nw + w];
            if (w < warp) { pa += va; pb += vb; }
            sa += va; sb += vb;
        }
        ea = pa + ia - a; eb = pb + ib - b; ta = sa; tb = sb;
        __syncthreads();
    }
};

__global__ void __launch_bounds__(REROOT_THREADS)
puct_reroot_kernel(const __grid_constant__ PuctEngine E, const int* __restrict__ rr_list,
                   const int* __restrict__ rr_count) {
    __shared__ int scan_sm[2 * (REROOT_THREADS / 32)];
    const int n = *rr_count;                       // a handful per step: one CTA each
    for (int k = blockIdx.x; k < n; k += gridDim.x) {
        const int g = rr_list[k];
        GameCtl& c = E.ctl[g];
        GameView V = gzp_view(&E, g);
        GzpCta grp;
        grp.sm = scan_sm;
        const int new_root = c.req_parent;
        __syncthreads();
        gzp_reroot_compact(V, new_root, (int)threadIdx.x, (int)blockDim.x, grp);
        __syncthreads();
        if (threadIdx.x == 0) c.req = REQ_NONE;
    }
}

__global__ void puct_init_kernel(const __grid_constant__ PuctEngine E, int mode) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= E.G) return;
    GameCtl& c = E.ctl[g];
    c.mode = mode;
    c.phase = mode == 0 ? PH_NEW_GAME : PH_IDLE;
    c.eval_node = -1;
    // evaluator stream then self-play stream, in game order: the construction order of the
    // reference (selfplaymanager.cpp:133-149) under oracle/shims/k273/rng.h seeding
    const uint64_t phi = 0x9E3779B97F4A7C15ULL;
    gzp_rng_seed(c.rng_eval, E.seed + (uint64_t)(2 * g) * phi);
    gzp_rng_seed(c.rng_sp, E.seed + (uint64_t)(2 * g + 1) * phi);
}

struct PuctTotals {
    unsigned long long evals, moves, games, samples, playouts;
    int errors, halted, first_error, max_nodes, max_children;
    int dupes, no_samples, fp_resign0, fp_resign1, early_run_to_ends, resigns, aborts;
    int waiting_eval;
    unsigned long long sel_nodes, sel_children, new_children;
    unsigned long long depth_sum;      // sum over games of the current root's game depth (mean ply = / G)
    unsigned long long transpositions; // expansions that attached an existing node (lookup_transpositions)
    int error_restarts;                // games abandoned at a per-game capacity limit (their slot started a new game)
};

__global__ void puct_totals_kernel(const __grid_constant__ PuctEngine E, PuctTotals* out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= E.G) return;
    const GameCtl& c = E.ctl[g];
    atomicAdd(&out->evals, (unsigned long long)c.total_evals);
    atomicAdd(&out->moves, (unsigned long long)c.total_moves);
    atomicAdd(&out->games, (unsigned long long)c.total_games);
    atomicAdd(&out->samples, (unsigned long long)c.total_samples);
    atomicAdd(&out->playouts, (unsigned long long)c.total_playouts);
    if (c.error) { atomicAdd(&out->errors, 1); atomicMax(&out->first_error, c.error); }
    if (c.phase == PH_HALT) atomicAdd(&out->halted, 1);
    atomicMax(&out->max_nodes, c.n_nodes);
    atomicMax(&out->max_children, c.n_children);
    atomicAdd(&out->dupes, c.saw_dupes);
    atomicAdd(&out->no_samples, c.no_samples_taken);
    atomicAdd(&out->fp_resign0, c.false_positive_resigns0);
    atomicAdd(&out->fp_resign1, c.false_positive_resigns1);
    atomicAdd(&out->early_run_to_ends, c.number_early_run_to_ends);
    atomicAdd(&out->resigns, c.number_resigns);
    atomicAdd(&out->aborts, c.number_aborts_game_length);
    if (c.eval_node >= 0) atomicAdd(&out->waiting_eval, 1);
    atomicAdd(&out->sel_nodes, (unsigned long long)c.total_sel_nodes);
    atomicAdd(&out->sel_children, (unsigned long long)c.total_sel_children);
    atomicAdd(&out->new_children, (unsigned long long)c.total_new_children);
    atomicAdd(&out->depth_sum, (unsigned long long)(c.game_depth > 0 ? c.game_depth : 0));
    atomicAdd(&out->error_restarts, c.error_restarts);
    atomicAdd(&out->transpositions, (unsigned long long)c.total_transpositions);
}

// ---- synthetic network (parity / bench utility): same function as oracle/ref_harness.cpp
template <typename T>
__device__ __forceinline__ float synth_load(const T* p, int64_t i);
template <>
__device__ __forceinline__ float synth_load<float>(const float* p, int64_t i) { return p[i]; }
template <>
__device__ __forceinline__ float synth_load<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) { return __bfloat162float(p[i]); }

template <typename T>
__global__ void synth_nn_kernel(const T* __restrict__ planes, int64_t n, int total, uint64_t seed, int n_roles,
                                const int* __restrict__ role_sizes_off /* [2*n_roles]: size, offset */,
                                int policy_stride, int n_rewards, float* __restrict__ policy,
                                float* __restrict__ value) {
    // one warp per row
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    uint64_t h = 0;
    for (int e = lane; e < total; e += 32) {
        const float v = synth_load<T>(planes, row * total + e);
        if (v != 0.0f) h += gz_synth_elem((uint32_t)e, __float_as_uint(v), seed);
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) h += __shfl_xor_sync(0xffffffffu, h, s);
    for (int r = 0; r < n_roles; r++) {
        const int sz = role_sizes_off[2 * r], off = role_sizes_off[2 * r + 1];
        for (int a = lane; a < sz; a += 32)
            policy[row * policy_stride + off + a] = gz_synth_policy(h, (uint32_t)r, (uint32_t)a);
    }
    if (lane == 0) {
        float out[8];
        gz_synth_values(h, n_rewards, out);
        for (int i = 0; i < n_rewards; i++) value[row * n_rewards + i] = out[i];
    }
}

// ------------------------------------------------------------------------------------------------
struct gz_puct {
    int device = 0;
    PuctEngine E{};
    void* d_arena = nullptr;
    size_t arena_bytes = 0;
    gz_propnet* pn = nullptr;
    gz_transformer* tr = nullptr;
    PuctTotals* d_totals = nullptr;
    int* d_role_sizes_off = nullptr;
    int mode = 0;
    int total_size = 0;
    int cache_words = 0;      // per-warp shared-memory words for the child cache (0: use the global scratch)
    int plane_c_pad = 0;      // 0: planes are NCHW rows; >0: channels-last rows padded to this channel count
    uint32_t* d_pn_scratch = nullptr;   // wide games: this engine's propnet value-array scratch
    int* d_rr_list = nullptr;           // ids of the games that asked for a re-root this step, and their count
    int* d_rr_count = nullptr;
    // re-rooting runs on a side stream, concurrently with the propnet pass, node creation and encoding
    // of the same step (the games that re-root take no part in those); joined before the step ends
    cudaStream_t rr_stream = nullptr;
    cudaEvent_t rr_fork = nullptr, rr_join = nullptr, gl_fork = nullptr;
    // games with a deep goal cone (reversi's piece counting: 596 chunks against 199 for everything else):
    // the main pass skips it; the terminal states are compacted and evaluated on the side stream, followed
    // by the creation of their (terminal) nodes — nothing on the way to the network needs them
    bool split_goals = false;
    bool side_pending = false;       // side-stream work of the last step has not been joined yet
    int *t_idx = nullptr, *t_count = nullptr;
    uint32_t *t_states = nullptr, *t_legal = nullptr;
    uint8_t *t_term = nullptr, *t_goals = nullptr;

    int* d_status = nullptr;         // gz_puct_status scratch
    // compacted network batches (gz_puct_compact_requests / gz_puct_advance_compact)
    int *c_rows = nullptr, *c_count = nullptr;     // waiting rows (ascending); [0] how many wait, [1] how many were handed out
    float *c_pol = nullptr, *c_val = nullptr;      // predictions scattered back to game rows
    uint8_t* c_answered = nullptr;                 // per game: its prediction arrived with this step
    struct HostPoll* hp = nullptr;   // host-buffer poll state (gz_puct_host_poll_*), created on first use
};

// State of the host-buffer poll (the reference's numpy Supervisor.poll contract): pinned hand-off
// buffers on the host, staging + compaction buffers on the device, one event per engine.
struct HostPoll {
    float *h_pol = nullptr, *h_val = nullptr, *h_planes = nullptr;   // pinned
    int* h_count = nullptr;                                          // pinned
    float *d_stage_pol = nullptr, *d_stage_val = nullptr;            // [G][..] predictions in hand-out order
    float *d_pol = nullptr, *d_val = nullptr;                        // [G][..] predictions in game order
    float *d_planes = nullptr, *d_planes_c = nullptr;                // [G][total] all games / (unused)
    float* h_planes_dev = nullptr;                                   // device address of the pinned h_planes
    int *d_rows = nullptr, *d_count = nullptr;
    cudaEvent_t ev = nullptr, ev_in = nullptr;
    int outstanding = 0;      // rows handed out by the last wait
    bool in_flight = false;
};

static void fill_puct_conf(PuctConf& p, const double* c) {
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

static void fill_sp_conf(SpConf& s, const double* c) {
    int i = 0;
    s.oscillate_sampling_pct = (float)c[i++]; s.temperature_for_policy = (float)c[i++];
    s.evals_per_move = (int)c[i++]; s.resign0_score_probability = (float)c[i++]; s.resign0_pct = (float)c[i++];
    s.resign1_score_probability = (float)c[i++]; s.resign1_pct = (float)c[i++]; s.run_to_end_pct = (float)c[i++];
    s.run_to_end_evals = (int)c[i++]; s.run_to_end_early_score = (float)c[i++];
    s.run_to_end_minimum_game_depth = (int)c[i++]; s.abort_max_length = (int)c[i++];
}

extern "C" int gz_puct_set_config(gz_puct* h, const double* sp_conf, const double* puct_conf, const double* rte_conf) {
    if (!h || !sp_conf || !puct_conf || !rte_conf) { gz_set_error("null argument"); return GZ_ERR_ARG; }
    fill_sp_conf(h->E.conf, sp_conf);
    fill_puct_conf(h->E.conf.puct, puct_conf);
    fill_puct_conf(h->E.conf.run_to_end, rte_conf);
    if ((h->E.conf.puct.lookup_transpositions || h->E.conf.run_to_end.lookup_transpositions) && h->d_arena && !h->E.tt_cap) {
        // the table is sized and allocated with the trees
        gz_set_error("lookup_transpositions must be set when the engine is created");
        return GZ_ERR_LIMIT;
    }
    return GZ_OK;
}

extern "C" int gz_puct_create(gz_propnet* pn, gz_transformer* tr, int n_games, int mode, int node_cap,
                              int child_cap, int max_pending_samples, int ready_cap, const double* sp_conf,
                              const double* puct_conf, const double* rte_conf, uint64_t seed, gz_puct** out) {
    if (!pn || !tr || !out || n_games <= 0 || node_cap < 16 || child_cap < 16) { gz_set_error("bad argument"); return GZ_ERR_ARG; }
    if (!tr->finalized) { gz_set_error("transformer must be finalised"); return GZ_ERR_ARG; }
    if (tr->num_prev_states > 1) {
        // the engine keeps ONE ancestor state per leaf and per sample (every shipped generation has
        // num_previous_states = 1); more would silently encode / sample different planes than
        // the reference (node.cpp:450-460, selfplaymanager.cpp:78-90)
        gz_set_error("the device engine supports num_previous_states <= 1, transformer has %d", tr->num_prev_states);
        return GZ_ERR_LIMIT;
    }
    *out = nullptr;
    gz_puct* h = new gz_puct();
    h->pn = pn;
    h->tr = tr;
    h->mode = mode;
    h->device = tr->device;
    PuctEngine& E = h->E;
    E.G = n_games;
    E.R = (int)gz_propnet_info(pn, GZ_INFO_N_ROLES);
    E.SW = (int)gz_propnet_info(pn, GZ_INFO_STATE_WORDS);
    E.LW = (int)gz_propnet_info(pn, GZ_INFO_LEGAL_WORDS);
    E.NCAP = node_cap;
    E.CCAP = child_cap;
    E.MAXD = 128;     // longest playout path (plies below the root)
    E.n_rewards = tr->num_rewards;
    int off = 0, maxa = 0;
    std::vector<int> rso;
    for (int r = 0; r < E.R; r++) {
        E.role_sizes[r] = (int)gz_propnet_info(pn, GZ_INFO_ROLE_SIZE0 + r);
        E.policy_off[r] = off;
        rso.push_back(E.role_sizes[r]);
        rso.push_back(off);
        off += E.role_sizes[r];
        maxa = std::max(maxa, E.role_sizes[r]);
    }
    for (int r = 0; r <= E.R; r++) E.legal_word_off[r] = (int)gz_propnet_info(pn, GZ_INFO_LEGAL_WORD_OFF0 + r);
    E.policy_stride = off;
    E.max_actions = maxa;
    E.MAXPEND = max_pending_samples > 0 ? max_pending_samples : 128;
    E.srec = gzp_sample_record_words(E.SW, maxa);
    E.READYCAP = ready_cap > 0 ? ready_cap : 1 << 16;
    E.uniq_cap = 1 << 20;
    E.max_iters_per_step = 16;
    // self-play: the slowest warp sets the duration of puct_pre_kernel for all games (end-game trees run playout
    // after playout into terminal nodes without ever needing the pipeline: 8 playouts = 210 us against 40 us for
    // a game that descends once). Budget in tree levels, forced moves included, plus 2 per finished playout;
    // measured on reversi (bench.py --max-work, steady state): 6 -> 13.95 K, 12 -> 14.06 K, 24 -> 13.29 K,
    // 40 -> 12.52 K moves/s. A single player tree has nobody to wait for.
    E.max_work_per_step = mode == 0 ? 12 : 1 << 20;
    // measurement switch: prefetching the node records of a node's expanded children during select was measured
    // to change nothing (0.186-0.194 ms either way: the kernel is bound by the instruction latency of its slowest
    // warps, not by that round trip) and costs 30 % more DRAM traffic, so it is off unless asked for
    { const char* pf = getenv("GZ_PUCT_PREFETCH"); E.prefetch_nodes = (pf && pf[0] == '1') ? 1 : 0; }
    E.dbg_cycles = nullptr;
    E.x_answered = nullptr;
    E.seed = seed;
    if ((int)tr->policy_sizes.size() != E.R) { gz_set_error("transformer policy heads != roles"); delete h; return GZ_ERR_ARG; }
    for (int r = 0; r < E.R; r++)
        if (tr->policy_sizes[r] != E.role_sizes[r]) { gz_set_error("policy size of role %d: %d != %d actions", r, tr->policy_sizes[r], E.role_sizes[r]); delete h; return GZ_ERR_ARG; }
    int rc = gz_puct_set_config(h, sp_conf, puct_conf, rte_conf);
    if (rc != GZ_OK) { delete h; return rc; }
    // evaluation rows per game: a Player's search hands up to PuctConfig.batch_size leaves of its tree to the network
    // at once (player.cpp:27-36: the scheduler's batch is the configuration's); self-play keeps one per game
    E.EB = 1;
    if (mode == GZ_PUCT_MODE_PLAYER && E.conf.puct.batch_size > 1) {
        if (E.conf.puct.batch_size > GZP_MAX_EB) {
            gz_set_error("PuctConfig.batch_size %d: the device player runs at most %d playout coroutines per tree", E.conf.puct.batch_size, GZP_MAX_EB);
            delete h;
            return GZ_ERR_LIMIT;
        }
        E.EB = E.conf.puct.batch_size;
    }
    E.tt_cap = 0;
    if (E.conf.puct.lookup_transpositions || E.conf.run_to_end.lookup_transpositions) {
        // transposition table (evaluator.cpp:52,144-173): load factor <= 1/2 of the node arena
        E.tt_cap = 32;
        while (E.tt_cap < 2 * E.NCAP) E.tt_cap *= 2;
    }
    GZ_CUDA(cudaSetDevice(h->device));
    h->arena_bytes = gzp_layout(&E, nullptr);
    cudaError_t e = cudaMalloc(&h->d_arena, h->arena_bytes);
    if (e != cudaSuccess) {
        gz_set_error("cudaMalloc of %.1f MB for %d game trees failed: %s", h->arena_bytes / 1e6, n_games, cudaGetErrorString(e));
        delete h;
        return GZ_ERR_CUDA;
    }
    GZ_CUDA(cudaMemset(h->d_arena, 0, h->arena_bytes));
    gzp_layout(&E, (uint8_t*)h->d_arena);
    // initial state + hash mask
    {
        std::vector<uint32_t> init(E.SW), mask(E.SW, 0);
        GZ_CUDA(cudaMemcpy(init.data(), gz_propnet_init_words(pn), E.SW * 4, cudaMemcpyDeviceToHost));
        for (int b : tr->board_base) mask[b >> 5] |= 1u << (b & 31);
        for (int b : tr->ctrl_base) mask[b >> 5] |= 1u << (b & 31);
        GZ_CUDA(cudaMemcpy((void*)E.init_state, init.data(), E.SW * 4, cudaMemcpyHostToDevice));
        GZ_CUDA(cudaMemcpy((void*)E.uniq_mask, mask.data(), E.SW * 4, cudaMemcpyHostToDevice));
    }
    GZ_CUDA(cudaMalloc(&h->d_totals, sizeof(PuctTotals)));
    GZ_CUDA(cudaStreamCreateWithFlags(&h->rr_stream, cudaStreamNonBlocking));
    GZ_CUDA(cudaEventCreateWithFlags(&h->rr_fork, cudaEventDisableTiming));
    GZ_CUDA(cudaEventCreateWithFlags(&h->rr_join, cudaEventDisableTiming));
    GZ_CUDA(cudaEventCreateWithFlags(&h->gl_fork, cudaEventDisableTiming));
    if (gz_propnet_goal_chunks(pn) >= 64) {
        h->split_goals = true;
        GZ_CUDA(cudaMalloc((void**)&h->t_idx, (size_t)E.G * 4));
        GZ_CUDA(cudaMalloc((void**)&h->t_count, 64));
        GZ_CUDA(cudaMalloc((void**)&h->t_states, (size_t)E.G * E.SW * 4));
        GZ_CUDA(cudaMalloc((void**)&h->t_legal, (size_t)E.G * E.LW * 4));
        GZ_CUDA(cudaMalloc((void**)&h->t_term, (size_t)E.G));
        GZ_CUDA(cudaMalloc((void**)&h->t_goals, (size_t)E.G * E.R));
        GZ_CUDA(cudaMemset(h->t_count, 0, 64));
    }
    GZ_CUDA(cudaMalloc((void**)&h->d_rr_list, (size_t)E.G * 4));
    GZ_CUDA(cudaMalloc((void**)&h->d_rr_count, 64));
    GZ_CUDA(cudaMemset(h->d_rr_count, 0, 64));
    if (gz_propnet_wide_scratch_bytes(pn) > 0) {
        GZ_CUDA(cudaMalloc((void**)&h->d_pn_scratch, (size_t)gz_propnet_wide_scratch_bytes(pn)));
        GZ_CUDA(cudaMemset(h->d_pn_scratch, 0, (size_t)gz_propnet_wide_scratch_bytes(pn)));
    }
    GZ_CUDA(cudaMalloc(&h->d_role_sizes_off, rso.size() * sizeof(int)));
    GZ_CUDA(cudaMemcpy(h->d_role_sizes_off, rso.data(), rso.size() * sizeof(int), cudaMemcpyHostToDevice));
    h->total_size = tr->E.total;
    {
        const size_t words = gzp_cache_words(E.max_actions);
        const size_t bytes = (words + 2 * (size_t)E.MAXD) * 4 * PUCT_WARPS_PER_BLOCK;
        if (bytes <= 96 * 1024) {
            h->cache_words = (int)words;
            if (bytes > 48 * 1024)
                GZ_CUDA(cudaFuncSetAttribute(puct_pre_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        }
    }
    puct_init_kernel<<<(E.G + 127) / 128, 128>>>(E, mode);
    GZ_CUDA(cudaGetLastError());
    GZ_CUDA(cudaDeviceSynchronize());
    *out = h;
    return GZ_OK;
}

// Per-step limits of the controller (see PuctEngine): iterations and tree levels descended per game and step.
extern "C" int gz_puct_set_step_limits(gz_puct* h, int max_iters, int max_work) {
    if (!h || max_iters < 1 || max_work < 1) { gz_set_error("bad step limits"); return GZ_ERR_ARG; }
    h->E.max_iters_per_step = max_iters;
    h->E.max_work_per_step = max_work;
    return GZ_OK;
}

extern "C" void gz_puct_destroy(gz_puct* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->d_arena) cudaFree(h->d_arena);
    if (h->d_totals) cudaFree(h->d_totals);
    if (h->d_pn_scratch) cudaFree(h->d_pn_scratch);
    if (h->d_status) cudaFree(h->d_status);
    { void* cv[] = {h->c_rows, h->c_count, h->c_pol, h->c_val, h->c_answered};
      for (void* d : cv) if (d) cudaFree(d); }
    if (h->d_rr_list) cudaFree(h->d_rr_list);
    if (h->d_rr_count) cudaFree(h->d_rr_count);
    if (h->rr_fork) cudaEventDestroy(h->rr_fork);
    if (h->rr_join) cudaEventDestroy(h->rr_join);
    if (h->gl_fork) cudaEventDestroy(h->gl_fork);
    { void* tv[] = {h->t_idx, h->t_count, h->t_states, h->t_legal, h->t_term, h->t_goals};
      for (void* d : tv) if (d) cudaFree(d); }
    if (h->rr_stream) cudaStreamDestroy(h->rr_stream);
    if (h->d_role_sizes_off) cudaFree(h->d_role_sizes_off);
    if (h->hp) {
        HostPoll* q = h->hp;
        if (q->h_pol) cudaFreeHost(q->h_pol);
        if (q->h_val) cudaFreeHost(q->h_val);
        if (q->h_planes) cudaFreeHost(q->h_planes);
        if (q->h_count) cudaFreeHost(q->h_count);
        void* dv[] = {q->d_stage_pol, q->d_stage_val, q->d_pol, q->d_val, q->d_planes, q->d_planes_c, q->d_rows, q->d_count};
        for (void* d : dv) if (d) cudaFree(d);
        if (q->ev) cudaEventDestroy(q->ev);
        if (q->ev_in) cudaEventDestroy(q->ev_in);
        delete q;
    }
    delete h;
}

extern "C" int64_t gz_puct_info(const gz_puct* h, int what) {
    if (!h) return -1;
    switch (what) {
        case 0: return h->E.G;
        case 1: return h->total_size;
        case 2: return h->E.policy_stride;
        case 3: return h->E.n_rewards;
        case 4: return h->E.srec;
        case 5: return (int64_t)h->arena_bytes;
        case 6: return h->E.SW;
        case 7: return h->E.max_actions;
        case 8: return h->E.READYCAP;
        case 9: return h->split_goals ? 10 : 5;   // kernels of this library launched per engine step (sub_iters == 1)
        case 10: return h->E.EB;                  // evaluation rows per game (planes / policy / value rows = G * EB)
        default: return -1;
    }
}

// One engine step. d_policy [G][policy_stride] f32 and d_value [G][n_rewards] f32 are the network's
// outputs for the planes produced by the previous advance (ignored when nn_fresh == 0, e.g. on the
// very first call). d_planes receives [G][total_size] of plane_dtype, one row per game.
// ---- helpers of the forked goal pass
__global__ void term_gather_kernel(const uint32_t* __restrict__ states, const int* __restrict__ idx,
                                   const int* __restrict__ count, int SW, uint32_t* __restrict__ out) {
    const int n = *count;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)n * SW; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / SW), c = (int)(i % SW);
        out[i] = states[(int64_t)idx[r] * SW + c];
    }
}
__global__ void term_scatter_kernel(const uint8_t* __restrict__ t_goals, const int* __restrict__ idx,
                                    const int* __restrict__ count, int R, uint8_t* __restrict__ goals) {
    const int n = *count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * R; i += gridDim.x * blockDim.x)
        goals[(int64_t)idx[i / R] * R + (i % R)] = t_goals[i];
}
__global__ void __launch_bounds__(1024) hostpoll_compact_kernel(const uint8_t* __restrict__ need, int G,
                                                                int* __restrict__ rows, int* __restrict__ count);

// Main stream waits for the side-stream work (re-root, goal pass) of the last step.
static int join_side(gz_puct* h, cudaStream_t s) {
    if (h->side_pending) {
        GZ_CUDA(cudaStreamWaitEvent(s, h->rr_join, 0));
        h->side_pending = false;
    }
    return GZ_OK;
}

static int advance_impl(gz_puct* h, const float* d_policy, const float* d_value, int nn_fresh, int sub_iters,
                        void* d_planes, int plane_dtype, void* stream, cudaEvent_t* evs /* 6 or null */,
                        bool defer_join) {
    if (!h) { gz_set_error("null engine"); return GZ_ERR_ARG; }
    if (nn_fresh && (!d_policy || !d_value)) { gz_set_error("network outputs missing"); return GZ_ERR_ARG; }
    if (sub_iters < 1) sub_iters = 1;
    GZ_CUDA(cudaSetDevice(h->device));
    PuctEngine E = h->E;
    E.nn_policy = d_policy;
    E.nn_value = d_value;
    cudaStream_t s = (cudaStream_t)stream;
    const int blocks = (E.G + PUCT_WARPS_PER_BLOCK - 1) / PUCT_WARPS_PER_BLOCK;
    int rc = join_side(h, s);                       // a deferred join of the previous step
    if (rc != GZ_OK) return rc;
    if (evs) GZ_CUDA(cudaEventRecord(evs[0], s));
    for (int it = 0; it < sub_iters; it++) {
        GZ_CUDA(cudaMemsetAsync(h->d_rr_count, 0, 4, s));
        puct_pre_kernel<<<blocks, PUCT_WARPS_PER_BLOCK * 32,
                          h->cache_words ? ((size_t)h->cache_words + 2 * (size_t)E.MAXD) * 4 * PUCT_WARPS_PER_BLOCK : 0, s>>>(
            E, (it == 0 && nn_fresh) ? 1 : 0, h->cache_words, h->d_rr_list, h->d_rr_count);
        GZ_CUDA(cudaGetLastError());
        if (evs) GZ_CUDA(cudaEventRecord(evs[1], s));
        // the last iteration forks: re-rooting, and (split_goals) the goal cone of the terminal states with the
        // creation of their nodes, run on the side stream beside the propnet pass / node creation / encoding
        const bool fork = (evs == nullptr) && (it + 1 == sub_iters);
        const bool split = fork && h->split_goals;
        if (fork) {
            GZ_CUDA(cudaEventRecord(h->rr_fork, s));
            GZ_CUDA(cudaStreamWaitEvent(h->rr_stream, h->rr_fork, 0));
            puct_reroot_kernel<<<std::min(E.G, 148), REROOT_THREADS, 0, h->rr_stream>>>(E, h->d_rr_list, h->d_rr_count);
            GZ_CUDA(cudaGetLastError());
        } else {
            puct_reroot_kernel<<<std::min(E.G, 148), REROOT_THREADS, 0, s>>>(E, h->d_rr_list, h->d_rr_count);
            GZ_CUDA(cudaGetLastError());
        }
        if (evs) GZ_CUDA(cudaEventRecord(evs[2], s));
        rc = gz_sm_transition_scratch(h->pn, E.G, E.x_in_state, E.x_moves, E.x_state, E.x_legal, E.x_term, E.x_goals,
                                      split ? GZ_GOALS_NEVER_INTERNAL : GZ_GOALS_TERMINAL_ONLY, stream, h->d_pn_scratch);
        if (rc != GZ_OK) return rc;
        if (evs) GZ_CUDA(cudaEventRecord(evs[3], s));
        if (split) {
            cudaStream_t ss = h->rr_stream;
            GZ_CUDA(cudaEventRecord(h->gl_fork, s));
            GZ_CUDA(cudaStreamWaitEvent(ss, h->gl_fork, 0));
            hostpoll_compact_kernel<<<1, 1024, 0, ss>>>(E.x_term, E.G, h->t_idx, h->t_count);
            term_gather_kernel<<<std::min((E.G * E.SW + 255) / 256, 592), 256, 0, ss>>>(E.x_state, h->t_idx, h->t_count, E.SW, h->t_states);
            GZ_CUDA(cudaGetLastError());
            rc = gz_sm_eval_devcount(h->pn, E.G, h->t_count, h->t_states, h->t_legal, h->t_term, h->t_goals,
                                     GZ_GOALS_ALWAYS, (void*)ss, h->d_pn_scratch);
            if (rc != GZ_OK) return rc;
            term_scatter_kernel<<<std::min((E.G * E.R + 255) / 256, 148), 256, 0, ss>>>(h->t_goals, h->t_idx, h->t_count, E.R, E.x_goals);
            puct_post_kernel<<<blocks, PUCT_WARPS_PER_BLOCK * 32, 0, ss>>>(E, 2);       // nodes of the terminal states
            GZ_CUDA(cudaGetLastError());
            puct_post_kernel<<<blocks, PUCT_WARPS_PER_BLOCK * 32, 0, s>>>(E, 1);        // everything else
        } else {
            puct_post_kernel<<<blocks, PUCT_WARPS_PER_BLOCK * 32, 0, s>>>(E, 0);
        }
        GZ_CUDA(cudaGetLastError());
        if (fork) {
            GZ_CUDA(cudaEventRecord(h->rr_join, h->rr_stream));
            h->side_pending = true;
        }
        if (evs) GZ_CUDA(cudaEventRecord(evs[4], s));
    }
    if (d_planes) {
        // every row is encoded (rows of games that do not wait for the network hold a stale but valid
        // state; their network outputs are never read)
        rc = gz_encode_layout(h->tr, (int64_t)E.G * E.EB, E.SW, E.e_state, E.e_prev, E.e_prevcnt, 1, d_planes, plane_dtype,
                              h->plane_c_pad, stream);
        if (rc != GZ_OK) return rc;
    }
    if (!defer_join) { rc = join_side(h, s); if (rc != GZ_OK) return rc; }
    if (evs) GZ_CUDA(cudaEventRecord(evs[5], s));
    return GZ_OK;
}

extern "C" int gz_puct_advance(gz_puct* h, const float* d_policy, const float* d_value, int nn_fresh, int sub_iters,
                               void* d_planes, int plane_dtype, void* stream) {
    return advance_impl(h, d_policy, d_value, nn_fresh, sub_iters, d_planes, plane_dtype, stream, nullptr, false);
}

// Same step, but the side-stream work it forks (re-rooting, the goal pass of terminal states) is left
// running: the caller may enqueue the network forward right away and must call gz_puct_join (or the next
// gz_puct_advance*, which joins first) on the same stream before it reads or steps the engine again.
extern "C" int gz_puct_advance_async(gz_puct* h, const float* d_policy, const float* d_value, int nn_fresh,
                                     int sub_iters, void* d_planes, int plane_dtype, void* stream) {
    return advance_impl(h, d_policy, d_value, nn_fresh, sub_iters, d_planes, plane_dtype, stream, nullptr, true);
}
extern "C" int gz_puct_join(gz_puct* h, void* stream) {
    if (!h) { gz_set_error("null engine"); return GZ_ERR_ARG; }
    GZ_CUDA(cudaSetDevice(h->device));
    return join_side(h, (cudaStream_t)stream);
}

// Measurement aid: one engine step (sub_iters == 1) with CUDA events between its launches on `stream`;
// waits for the step and returns the device time of each launch in milliseconds:
// h_ms5 = [puct_pre_kernel, puct_reroot_kernel, pn_kernel, puct_post_kernel, encode_kernel].
extern "C" int gz_puct_advance_timed(gz_puct* h, const float* d_policy, const float* d_value, int nn_fresh,
                                     void* d_planes, int plane_dtype, void* stream, float* h_ms5) {
    if (!h || !h_ms5) { gz_set_error("null argument"); return GZ_ERR_ARG; }
    GZ_CUDA(cudaSetDevice(h->device));
    cudaEvent_t evs[6];
    for (int i = 0; i < 6; i++) GZ_CUDA(cudaEventCreate(&evs[i]));
    int rc = advance_impl(h, d_policy, d_value, nn_fresh, 1, d_planes, plane_dtype, stream, evs, false);
    if (rc == GZ_OK) {
        cudaError_t e = cudaEventSynchronize(evs[5]);
        if (e != cudaSuccess) { gz_set_error("timed step failed: %s", cudaGetErrorString(e)); rc = GZ_ERR_CUDA; }
        for (int i = 0; i < 5 && rc == GZ_OK; i++) cudaEventElapsedTime(&h_ms5[i], evs[i], evs[i + 1]);
    }
    for (int i = 0; i < 6; i++) cudaEventDestroy(evs[i]);
    return rc;
}

// Planes layout written by gz_puct_advance: c_pad == 0 -> [G][C][H][W]; c_pad > 0 (multiple of 8) ->
// channels-last [G][H][W][c_pad] with zero planes beyond C.
extern "C" int gz_puct_set_plane_layout(gz_puct* h, int c_pad) {
    if (!h || c_pad < 0 || (c_pad & 7) || (c_pad && c_pad < h->tr->E.total / h->tr->E.channel_size)) { gz_set_error("bad plane layout"); return GZ_ERR_ARG; }
    h->plane_c_pad = c_pad;
    return GZ_OK;
}

// Measurement aid: d_buf = device int64 [G][4] or null; every puct_pre_kernel launch then records, per game, the
// cycles its warp spent, entry / exit phase, playouts finished and moves played in that step.
extern "C" int gz_puct_set_debug_cycles(gz_puct* h, long long* d_buf) {
    if (!h) { gz_set_error("null engine"); return GZ_ERR_ARG; }
    h->E.dbg_cycles = d_buf;
    return GZ_OK;
}

// ---- compacted network batches on the device ----------------------------------------------------
// In a long-running job a quarter of the games do not wait for the network after a step (their playouts ended in
// terminal or finalised nodes, their searches are over): the network need not see their rows. compact_requests
// lists the waiting rows (ascending), keeps at most `cap` of them and gathers their planes into a dense batch;
// advance_compact takes the predictions in that order. Requests beyond the cap simply stay pending — their games
// wait one more step — so a caller may size the batch from a STALE count (read back asynchronously, no host
// synchronisation in the loop) and still be correct.
__global__ void compact_cap_kernel(const int* __restrict__ count, int cap, int* __restrict__ n_out) {
    *n_out = *count < cap ? *count : cap;
}
template <typename T>
__global__ void compact_gather_kernel(const T* __restrict__ planes, const int* __restrict__ rows,
                                      const int* __restrict__ n_ptr, int row_vec, T* __restrict__ out) {
    const int n = *n_ptr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)n * row_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / row_vec), c = (int)(i - (int64_t)r * row_vec);
        out[i] = planes[(int64_t)rows[r] * row_vec + c];
    }
}
__global__ void compact_scatter_kernel(const float* __restrict__ pol_c, const float* __restrict__ val_c,
                                       const int* __restrict__ rows, const int* __restrict__ n_ptr, int stride, int nrew,
                                       float* __restrict__ pol, float* __restrict__ val, uint8_t* __restrict__ answered) {
    const int n = *n_ptr;
    const int per = stride + nrew;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)n * per; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / per), c = (int)(i - (int64_t)r * per);
        const int g = rows[r];
        if (c < stride) pol[(int64_t)g * stride + c] = pol_c[(int64_t)r * stride + c];
        else val[(int64_t)g * nrew + (c - stride)] = val_c[(int64_t)r * nrew + (c - stride)];
        if (c == 0) answered[g] = 1;
    }
}

static int compact_init(gz_puct* h) {
    if (h->c_rows) return GZ_OK;
    if (h->E.EB != 1) { gz_set_error("compacted batches are a self-play feature (one evaluation row per game)"); return GZ_ERR_ARG; }
    const size_t G = (size_t)h->E.G;
    GZ_CUDA(cudaSetDevice(h->device));
    GZ_CUDA(cudaMalloc((void**)&h->c_rows, G * 4));
    GZ_CUDA(cudaMalloc((void**)&h->c_count, 256));
    GZ_CUDA(cudaMemset(h->c_count, 0, 256));
    GZ_CUDA(cudaMalloc((void**)&h->c_pol, G * h->E.policy_stride * 4));
    GZ_CUDA(cudaMalloc((void**)&h->c_val, G * h->E.n_rewards * 4));
    GZ_CUDA(cudaMalloc((void**)&h->c_answered, G));
    GZ_CUDA(cudaMemset(h->c_pol, 0, G * h->E.policy_stride * 4));
    GZ_CUDA(cudaMemset(h->c_val, 0, G * h->E.n_rewards * 4));
    GZ_CUDA(cudaMemset(h->c_answered, 0, G));
    return GZ_OK;
}

// d_planes: what gz_puct_advance wrote ([G] rows of row_bytes bytes, a multiple of 16); d_planes_c receives the
// first min(waiting, cap) waiting rows. h_count_pinned (optional, pinned host memory) receives the UNCAPPED number
// of waiting rows with an asynchronous copy.
extern "C" int gz_puct_compact_requests(gz_puct* h, int cap, const void* d_planes, int64_t row_bytes, void* d_planes_c,
                                        int* h_count_pinned, void* stream) {
    if (!h || !d_planes || !d_planes_c || cap < 1 || row_bytes <= 0 || (row_bytes & 1)) { gz_set_error("bad argument"); return GZ_ERR_ARG; }
    int rc = compact_init(h);
    if (rc != GZ_OK) return rc;
    GZ_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    // no join with the step's side-stream work here (re-rooting; goal values and nodes of terminal states): a terminal
    // state's node never asks for the network, so the request flags are complete once the main stream's node creation
    // has run, and the network forward that follows keeps overlapping that work
    hostpoll_compact_kernel<<<1, 1024, 0, s>>>(h->E.x_need_eval, h->E.G, h->c_rows, h->c_count);
    compact_cap_kernel<<<1, 1, 0, s>>>(h->c_count, cap, h->c_count + 1);
    // rows are copied in the widest unit their size allows (hexLG13: 5 x 13 x 13 bf16 = 1690 bytes -> 2-byte units)
    const int unit = (row_bytes % 16 == 0 && ((uintptr_t)d_planes % 16 == 0) && ((uintptr_t)d_planes_c % 16 == 0)) ? 16 : (row_bytes % 4 == 0 ? 4 : 2);
    const int row_vec = (int)(row_bytes / unit);
    const int64_t work = (int64_t)std::min(cap, h->E.G) * row_vec;
    const int ggrid = (int)std::min<int64_t>((work + 255) / 256, 1184);
    if (unit == 16)
        compact_gather_kernel<uint4><<<ggrid, 256, 0, s>>>((const uint4*)d_planes, h->c_rows, h->c_count + 1, row_vec, (uint4*)d_planes_c);
    else if (unit == 4)
        compact_gather_kernel<uint32_t><<<ggrid, 256, 0, s>>>((const uint32_t*)d_planes, h->c_rows, h->c_count + 1, row_vec, (uint32_t*)d_planes_c);
    else
        compact_gather_kernel<uint16_t><<<ggrid, 256, 0, s>>>((const uint16_t*)d_planes, h->c_rows, h->c_count + 1, row_vec, (uint16_t*)d_planes_c);
    GZ_CUDA(cudaGetLastError());
    if (h_count_pinned) GZ_CUDA(cudaMemcpyAsync(h_count_pinned, h->c_count, 4, cudaMemcpyDeviceToHost, s));
    return GZ_OK;
}

// The step that follows gz_puct_compact_requests: d_policy_c / d_value_c hold the network's outputs for the rows of
// the compacted batch, in its order.
extern "C" int gz_puct_advance_compact(gz_puct* h, const float* d_policy_c, const float* d_value_c, int sub_iters,
                                       void* d_planes, int plane_dtype, void* stream, int defer_join) {
    if (!h || !d_policy_c || !d_value_c) { gz_set_error("null argument"); return GZ_ERR_ARG; }
    int rc = compact_init(h);
    if (rc != GZ_OK) return rc;
    GZ_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    rc = join_side(h, s);
    if (rc != GZ_OK) return rc;
    const int64_t work = (int64_t)h->E.G * (h->E.policy_stride + h->E.n_rewards);
    compact_scatter_kernel<<<(int)std::min<int64_t>((work + 255) / 256, 1184), 256, 0, s>>>(
        d_policy_c, d_value_c, h->c_rows, h->c_count + 1, h->E.policy_stride, h->E.n_rewards, h->c_pol, h->c_val, h->c_answered);
    GZ_CUDA(cudaGetLastError());
    h->E.x_answered = h->c_answered;
    rc = advance_impl(h, h->c_pol, h->c_val, 1, sub_iters, d_planes, plane_dtype, stream, nullptr, defer_join != 0);
    h->E.x_answered = nullptr;
    return rc;
}

extern "C" const uint8_t* gz_puct_need_eval_ptr(const gz_puct* h) { return h ? h->E.x_need_eval : nullptr; }

// Copies the per-game "waiting for the network" flags (uint8 [G]) into a caller-owned device buffer.
extern "C" int gz_puct_copy_need_eval(gz_puct* h, uint8_t* d_out, void* stream) {
    if (!h || !d_out) { gz_set_error("null argument"); return GZ_ERR_ARG; }
    GZ_CUDA(cudaSetDevice(h->device));
    { int jrc = join_side(h, (cudaStream_t)stream); if (jrc != GZ_OK) return jrc; }
    GZ_CUDA(cudaMemcpyAsync(d_out, h->E.x_need_eval, (size_t)h->E.G * h->E.EB, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return GZ_OK;
}

// ---- host-buffer poll ---------------------------------------------------------------------------
// The reference's Supervisor.poll contract (src/cpp/pyobjects/common.cpp:160-218, src/cpp/supervisor.cpp:
// 101-168) with HOST buffers on both sides: predictions arrive as R policy arrays [n][A_r] + one value
// array [n][rewards] in the order the rows were handed out; the planes of the games that now wait for
// the network leave as [n_req][total] float32 in pinned memory owned by the engine.
// submit() stages the predictions and enqueues copy-in, scatter, the engine step, compaction, gather
// and copy-out on `stream` without waiting; wait() blocks on the engine's event. Two engines polled
// alternately overlap one engine's step with the caller's network call on the other's batch (the
// reference's two managers per worker, supervisor.cpp:79-99).
// staged predictions: role r's array [n][A_r] sits contiguously at sp + G * policy_off[r] (as it arrived)
struct HostPollRoles { int n_roles; int size[GZP_MAX_ROLES]; int off[GZP_MAX_ROLES]; };
__global__ void hostpoll_scatter_kernel(const float* __restrict__ sp, const float* __restrict__ sv,
                                        const int* __restrict__ rows, int n, int G, int stride, int nrew,
                                        const HostPollRoles rl, float* __restrict__ pol, float* __restrict__ val) {
    const int per = stride + nrew;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)n * per; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / per), c = (int)(i % per);
        const int g = rows[r];
        if (c < stride) {
            int k = 0;
            while (k + 1 < rl.n_roles && c >= rl.off[k + 1]) k++;
            pol[(int64_t)g * stride + c] = sp[(int64_t)G * rl.off[k] + (int64_t)r * rl.size[k] + (c - rl.off[k])];
        } else {
            val[(int64_t)g * nrew + (c - stride)] = sv[(int64_t)r * nrew + (c - stride)];
        }
    }
}

// single CTA: rows[] = indices of games with need_eval set, ascending; *count = how many
__global__ void __launch_bounds__(1024) hostpoll_compact_kernel(const uint8_t* __restrict__ need, int G,
                                                                int* __restrict__ rows, int* __restrict__ count) {
    __shared__ int wsum[32];
    __shared__ int base_sm;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base_sm = 0;
    __syncthreads();
    for (int g0 = 0; g0 < G; g0 += 1024) {
        const int g = g0 + threadIdx.x;
        const bool f = g < G && need[g] != 0;
        const unsigned b = __ballot_sync(0xffffffffu, f);
        if (lane == 0) wsum[warp] = __popc(b);
        __syncthreads();
        int pre = 0, tot = 0;
        for (int w = 0; w < 32; w++) { const int v = wsum[w]; if (w < warp) pre += v; tot += v; }
        const int base = base_sm;
        if (f) rows[base + pre + __popc(b & ((1u << lane) - 1))] = g;
        __syncthreads();
        if (threadIdx.x == 0) base_sm = base + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base_sm;
}

__global__ void hostpoll_gather_kernel(const float* __restrict__ planes, const int* __restrict__ rows,
                                       const int* __restrict__ count, int total, float* __restrict__ out) {
    const int n = *count;
    if ((total & 3) == 0) {
        const int t4 = total >> 2;
        const float4* src = reinterpret_cast<const float4*>(planes);
        float4* dst = reinterpret_cast<float4*>(out);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)n * t4; i += (int64_t)gridDim.x * blockDim.x) {
            const int r = (int)(i / t4), c = (int)(i % t4);
            dst[i] = src[(int64_t)rows[r] * t4 + c];
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)n * total; i += (int64_t)gridDim.x * blockDim.x) {
            const int r = (int)(i / total), c = (int)(i % total);
            out[i] = planes[(int64_t)rows[r] * total + c];
        }
    }
}

static int hostpoll_init(gz_puct* h) {
    if (h->hp) return GZ_OK;
    if (h->plane_c_pad) { gz_set_error("host polls need the reference plane layout (plane_c_pad == 0)"); return GZ_ERR_ARG; }
    HostPoll* q = new HostPoll();
    const size_t G = (size_t)h->E.G, st = (size_t)h->E.policy_stride, nr = (size_t)h->E.n_rewards, tot = (size_t)h->total_size;
    GZ_CUDA(cudaSetDevice(h->device));
    GZ_CUDA(cudaHostAlloc((void**)&q->h_pol, G * st * 4, cudaHostAllocDefault));
    GZ_CUDA(cudaHostAlloc((void**)&q->h_val, G * nr * 4, cudaHostAllocDefault));
    GZ_CUDA(cudaHostAlloc((void**)&q->h_planes, G * tot * 4, cudaHostAllocMapped));
    GZ_CUDA(cudaHostAlloc((void**)&q->h_count, 64, cudaHostAllocDefault));
    GZ_CUDA(cudaMalloc((void**)&q->d_stage_pol, G * st * 4));
    GZ_CUDA(cudaMalloc((void**)&q->d_stage_val, G * nr * 4));
    GZ_CUDA(cudaMalloc((void**)&q->d_pol, G * st * 4));
    GZ_CUDA(cudaMalloc((void**)&q->d_val, G * nr * 4));
    GZ_CUDA(cudaMalloc((void**)&q->d_planes, G * tot * 4));
    GZ_CUDA(cudaHostGetDevicePointer((void**)&q->h_planes_dev, q->h_planes, 0));
    GZ_CUDA(cudaMalloc((void**)&q->d_rows, G * 4));
    GZ_CUDA(cudaMalloc((void**)&q->d_count, 64));
    GZ_CUDA(cudaMemset(q->d_pol, 0, G * st * 4));
    GZ_CUDA(cudaMemset(q->d_val, 0, G * nr * 4));
    GZ_CUDA(cudaMemset(q->d_rows, 0, G * 4));
    GZ_CUDA(cudaMemset(q->d_count, 0, 64));
    GZ_CUDA(cudaEventCreateWithFlags(&q->ev, cudaEventDisableTiming));
    GZ_CUDA(cudaEventCreateWithFlags(&q->ev_in, cudaEventDisableTiming));
    q->h_count[0] = 0;
    h->hp = q;
    return GZ_OK;
}

extern "C" int gz_puct_host_poll_submit(gz_puct* h, int n_pred, const float* const* h_arrays, int n_arrays, void* stream) {
    if (!h) { gz_set_error("null engine"); return GZ_ERR_ARG; }
    int rc = hostpoll_init(h);
    if (rc != GZ_OK) return rc;
    HostPoll* q = h->hp;
    const PuctEngine& E = h->E;
    if (q->in_flight) { gz_set_error("host poll already submitted; call gz_puct_host_poll_wait first"); return GZ_ERR_ARG; }
    if (n_pred != q->outstanding) { gz_set_error("expected %d predictions, got %d", q->outstanding, n_pred); return GZ_ERR_ARG; }
    if (n_pred > 0 && (!h_arrays || n_arrays != E.R + 1)) { gz_set_error("expected %d policy arrays and one value array", E.R); return GZ_ERR_ARG; }
    GZ_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int st = E.policy_stride, nr = E.n_rewards;
    bool direct = false;
    if (n_pred > 0) {
        for (int r = 0; r <= E.R; r++)
            if (!h_arrays[r]) { gz_set_error("null prediction array"); return GZ_ERR_ARG; }
        // arrays that already live in pinned host memory are copied from where they are (one contiguous
        // copy per array); pageable arrays go through the engine's pinned staging rows first
        direct = true;
        for (int r = 0; r <= E.R && direct; r++) {
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, h_arrays[r]) != cudaSuccess || at.type != cudaMemoryTypeHost) direct = false;
        }
        cudaGetLastError();
        for (int r = 0; r < E.R; r++) {
            const size_t bytes = (size_t)n_pred * E.role_sizes[r] * 4;
            const size_t blk = (size_t)E.G * E.policy_off[r];
            const float* src = h_arrays[r];
            if (!direct) { memcpy(q->h_pol + blk, h_arrays[r], bytes); src = q->h_pol + blk; }
            GZ_CUDA(cudaMemcpyAsync(q->d_stage_pol + blk, src, bytes, cudaMemcpyHostToDevice, s));
        }
        {
            const float* src = h_arrays[E.R];
            if (!direct) { memcpy(q->h_val, h_arrays[E.R], (size_t)n_pred * nr * 4); src = q->h_val; }
            GZ_CUDA(cudaMemcpyAsync(q->d_stage_val, src, (size_t)n_pred * nr * 4, cudaMemcpyHostToDevice, s));
        }
        if (direct) GZ_CUDA(cudaEventRecord(q->ev_in, s));
        HostPollRoles rl;
        rl.n_roles = E.R;
        for (int r = 0; r < E.R; r++) { rl.size[r] = E.role_sizes[r]; rl.off[r] = E.policy_off[r]; }
        const int64_t work = (int64_t)n_pred * (st + nr);
        hostpoll_scatter_kernel<<<(int)std::min<int64_t>((work + 255) / 256, 1184), 256, 0, s>>>(
            q->d_stage_pol, q->d_stage_val, q->d_rows, n_pred, E.G, st, nr, rl, q->d_pol, q->d_val);
        GZ_CUDA(cudaGetLastError());
    }
    rc = gz_puct_advance(h, q->d_pol, q->d_val, n_pred > 0 ? 1 : 0, 1, q->d_planes, GZ_DTYPE_F32, stream);
    if (rc != GZ_OK) return rc;
    hostpoll_compact_kernel<<<1, 1024, 0, s>>>(E.x_need_eval, E.G, q->d_rows, q->d_count);
    GZ_CUDA(cudaGetLastError());
    const int64_t gw = (int64_t)E.G * h->total_size / 4;
    // the gather writes straight into the pinned host buffer (unified addressing: the kernel's 16-byte stores are
    // posted PCIe writes), so only the rows that wait for the network cross the bus — no device staging copy and no
    // device-to-host copy of all G rows whose length the host would have to know first
    hostpoll_gather_kernel<<<(int)std::min<int64_t>((gw + 255) / 256, 1184), 256, 0, s>>>(
        q->d_planes, q->d_rows, q->d_count, h->total_size, q->h_planes_dev);
    GZ_CUDA(cudaGetLastError());
    GZ_CUDA(cudaMemcpyAsync(q->h_count, q->d_count, 4, cudaMemcpyDeviceToHost, s));
    GZ_CUDA(cudaEventRecord(q->ev, s));
    q->outstanding = 0;
    q->in_flight = true;
    // the caller may reuse its arrays as soon as this returns (the reference copies them: supervisor.cpp:104-120)
    if (direct) GZ_CUDA(cudaEventSynchronize(q->ev_in));
    return GZ_OK;
}

// Blocks until the submitted poll has finished. *h_planes = pinned [n_req][total] float32 (valid until
// the next submit on this engine); returns n_req (>= 0) or a negative gz_status.
extern "C" int gz_puct_host_poll_wait(gz_puct* h, const float** h_planes) {
    if (!h || !h->hp || !h->hp->in_flight) { gz_set_error("no host poll in flight"); return GZ_ERR_ARG; }
    HostPoll* q = h->hp;
    GZ_CUDA(cudaEventSynchronize(q->ev));
    q->in_flight = false;
    q->outstanding = q->h_count[0];
    if (h_planes) *h_planes = q->h_planes;
    return q->outstanding;
}

// totals: [evals, moves, games, samples, playouts, errors, halted, first_error, max_nodes, max_children,
//          dupes, no_samples, fp_resign0, fp_resign1, early_run_to_ends, resigns, aborts, waiting_eval,
//          ready_count, select_nodes, select_children, new_children, depth_sum, dropped_samples, error_restarts,
//          transpositions] (26 doubles)
extern "C" int gz_puct_totals(gz_puct* h, double* out, void* stream) {
    if (!h || !out) { gz_set_error("null argument"); return GZ_ERR_ARG; }
    GZ_CUDA(cudaSetDevice(h->device));
    { int jrc = join_side(h, (cudaStream_t)stream); if (jrc != GZ_OK) return jrc; }
    cudaStream_t s = (cudaStream_t)stream;
    GZ_CUDA(cudaMemsetAsync(h->d_totals, 0, sizeof(PuctTotals), s));
    puct_totals_kernel<<<(h->E.G + 127) / 128, 128, 0, s>>>(h->E, h->d_totals);
    GZ_CUDA(cudaGetLastError());
    PuctTotals t;
    unsigned long long ready[2] = {0, 0};
    GZ_CUDA(cudaMemcpyAsync(&t, h->d_totals, sizeof(t), cudaMemcpyDeviceToHost, s));
    GZ_CUDA(cudaMemcpyAsync(ready, h->E.ready_count, 16, cudaMemcpyDeviceToHost, s));
    GZ_CUDA(cudaStreamSynchronize(s));
    out[0] = (double)t.evals; out[1] = (double)t.moves; out[2] = (double)t.games; out[3] = (double)t.samples;
    out[4] = (double)t.playouts; out[5] = t.errors; out[6] = t.halted; out[7] = t.first_error;
    out[8] = t.max_nodes; out[9] = t.max_children; out[10] = t.dupes; out[11] = t.no_samples;
    out[12] = t.fp_resign0; out[13] = t.fp_resign1; out[14] = t.early_run_to_ends; out[15] = t.resigns;
    out[16] = t.aborts; out[17] = t.waiting_eval;
    out[18] = (double)std::min<unsigned long long>(ready[0], (unsigned long long)h->E.READYCAP);
    out[19] = (double)t.sel_nodes; out[20] = (double)t.sel_children; out[21] = (double)t.new_children;
    out[22] = (double)t.depth_sum;
    out[23] = (double)ready[1];          // finished samples dropped because the ready buffer was full
    out[24] = t.error_restarts;
    out[25] = (double)t.transpositions;
    return GZ_OK;
}

// Copies the finished samples to the host (records of gz_puct_info(h,4) int32 words, layout SR_* in
// gz_puct_core.h) and empties the device buffer. Returns the number of records copied (<= cap), or
// a negative status. Synchronises `stream`.
// moves words [from, from + count) of buf to [0, count), from >= 0: one CTA, chunk by chunk, read before write
__global__ void __launch_bounds__(1024) ready_shift_kernel(int32_t* buf, long long from, long long count) {
    for (long long base = 0; base < count; base += blockDim.x) {
        const long long i = base + threadIdx.x;
        int32_t v = 0;
        if (i < count) v = buf[from + i];
        __syncthreads();
        if (i < count) buf[i] = v;
        __syncthreads();
    }
}

extern "C" int64_t gz_puct_fetch_samples(gz_puct* h, int32_t* h_out, int64_t cap, void* stream) {
    if (!h || (!h_out && cap > 0)) { gz_set_error("null argument"); return GZ_ERR_ARG; }
    cudaSetDevice(h->device);
    if (join_side(h, (cudaStream_t)stream) != GZ_OK) return GZ_ERR_CUDA;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long ready = 0;
    if (cudaMemcpyAsync(&ready, h->E.ready_count, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) { gz_set_error("sample count copy failed"); return GZ_ERR_CUDA; }
    const int64_t have = (int64_t)std::min<unsigned long long>(ready, (unsigned long long)h->E.READYCAP);
    const int64_t n = std::min<int64_t>(have, cap);
    if (n > 0 && cudaMemcpyAsync(h_out, h->E.ready, (size_t)n * h->E.srec * 4, cudaMemcpyDeviceToHost, s) != cudaSuccess) {
        gz_set_error("sample copy failed");
        return GZ_ERR_CUDA;
    }
    // records that did not fit into the caller's buffer stay, moved to the front, for the next fetch
    const unsigned long long left = (unsigned long long)(have - n);
    if (left > 0) {
        ready_shift_kernel<<<1, 1024, 0, s>>>(h->E.ready, (long long)n * h->E.srec, (long long)left * h->E.srec);
        if (cudaGetLastError() != cudaSuccess) { gz_set_error("sample shift failed"); return GZ_ERR_CUDA; }
    }
    cudaMemcpyAsync(h->E.ready_count, &left, 8, cudaMemcpyHostToDevice, s);
    cudaStreamSynchronize(s);
    return n;
}

// Measurement aid (bench.py): the NEXT game of every slot starts from d_states[g] (device, [G][state_words]) instead
// of the game's initial state, at game depth 0 — slots then sit at different plies from the first step on, which
// is the distribution a long-running self-play job is in. Positions must not be terminal.
__global__ void puct_start_pending_kernel(const __grid_constant__ PuctEngine E) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < E.G) E.ctl[g].start_pending = 1;
}
extern "C" int gz_puct_set_start_states(gz_puct* h, const uint32_t* d_states, void* stream) {
    if (!h || !d_states) { gz_set_error("null argument"); return GZ_ERR_ARG; }
    if (h->mode != 0) { gz_set_error("start states apply to self-play engines"); return GZ_ERR_ARG; }
    GZ_CUDA(cudaSetDevice(h->device));
    { int jrc = join_side(h, (cudaStream_t)stream); if (jrc != GZ_OK) return jrc; }
    GZ_CUDA(cudaMemcpyAsync(h->E.start_state, d_states, (size_t)h->E.G * h->E.SW * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    puct_start_pending_kernel<<<(h->E.G + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->E);
    GZ_CUDA(cudaGetLastError());
    return GZ_OK;
}

extern "C" int gz_puct_clear_unique_states(gz_puct* h, void* stream) {
    if (!h) return GZ_ERR_ARG;
    GZ_CUDA(cudaSetDevice(h->device));
    { int jrc = join_side(h, (cudaStream_t)stream); if (jrc != GZ_OK) return jrc; }
    GZ_CUDA(cudaMemsetAsync(h->E.uniq_keys, 0, (size_t)h->E.uniq_cap * 8, (cudaStream_t)stream));
    GZ_CUDA(cudaMemsetAsync(h->E.uniq_counts, 0, (size_t)h->E.uniq_cap * 4, (cudaStream_t)stream));
    return GZ_OK;
}

// ---- player mode (mode == 1): commands for game `g`; they take effect in the next advance.
__global__ void puct_cmd_kernel(const __grid_constant__ PuctEngine E, int g, int cmd, int arg, int evals) {
    GameCtl& c = E.ctl[g];
    c.cmd = cmd; c.cmd_arg = arg; c.cmd_evals = evals; c.cmd_done = 0;
}

// Player mode: the host's clock ran out (mode: GZP_STOP_NOW / GZP_STOP_IF_CONVERGED); the running search of game g
// ends at its next playout boundary. Status: [cmd_done, phase, error, waiting for the network] of game g (one
// 16-byte copy; synchronises `stream`).
__global__ void puct_stop_kernel(const __grid_constant__ PuctEngine E, int g, int mode) { E.ctl[g].stop_request = mode; }
extern "C" int gz_puct_request_stop(gz_puct* h, int g, int mode, void* stream) {
    if (!h || g < 0 || g >= h->E.G || mode < 0 || mode > 2) { gz_set_error("bad argument"); return GZ_ERR_ARG; }
    GZ_CUDA(cudaSetDevice(h->device));
    { int jrc = join_side(h, (cudaStream_t)stream); if (jrc != GZ_OK) return jrc; }
    puct_stop_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(h->E, g, mode);
    GZ_CUDA(cudaGetLastError());
    return GZ_OK;
}
__global__ void puct_status_kernel(const __grid_constant__ PuctEngine E, int g, int* out) {
    const GameCtl& c = E.ctl[g];
    int waiting = 0;                     // evaluation rows of this game handed to the network (rows g*EB .. + waiting)
    for (int k = 0; k < E.EB; k++) waiting += E.x_need_eval[(int64_t)g * E.EB + k] != 0;
    out[0] = c.cmd_done; out[1] = c.phase; out[2] = c.error; out[3] = waiting;
}
extern "C" int gz_puct_status(gz_puct* h, int g, int* h_out4, void* stream) {
    if (!h || !h_out4 || g < 0 || g >= h->E.G) { gz_set_error("bad argument"); return GZ_ERR_ARG; }
    GZ_CUDA(cudaSetDevice(h->device));
    { int jrc = join_side(h, (cudaStream_t)stream); if (jrc != GZ_OK) return jrc; }
    if (!h->d_status) GZ_CUDA(cudaMalloc((void**)&h->d_status, 64));
    puct_status_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(h->E, g, h->d_status);
    GZ_CUDA(cudaGetLastError());
    GZ_CUDA(cudaMemcpyAsync(h_out4, h->d_status, 16, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    GZ_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return GZ_OK;
}

extern "C" int gz_puct_command(gz_puct* h, int g, int cmd, int arg, int evals, const uint32_t* h_state, void* stream) {
    if (!h || g < 0 || g >= h->E.G) { gz_set_error("bad game index"); return GZ_ERR_ARG; }
    GZ_CUDA(cudaSetDevice(h->device));
    { int jrc = join_side(h, (cudaStream_t)stream); if (jrc != GZ_OK) return jrc; }
    cudaStream_t s = (cudaStream_t)stream;
    if (cmd == CMD_RESET) {
        if (h_state) GZ_CUDA(cudaMemcpyAsync(h->E.x_in_state + (size_t)g * h->E.SW, h_state, h->E.SW * 4, cudaMemcpyHostToDevice, s));
        else GZ_CUDA(cudaMemcpyAsync(h->E.x_in_state + (size_t)g * h->E.SW, h->E.init_state, h->E.SW * 4, cudaMemcpyDeviceToDevice, s));
    }
    puct_cmd_kernel<<<1, 1, 0, s>>>(h->E, g, cmd, arg, evals);
    GZ_CUDA(cudaGetLastError());
    return GZ_OK;
}

// Snapshot of game g: out_i[0..15] (out_i[9+r] = role r's only action at the root) = cmd_done, phase, error, choice, n_nodes, n_children, root visits,
// lead role; out_f[0] = root score for the lead role, out_f[1 + r] = score of the CHOSEN child's node for role r (-1:
// unexpanded / nothing chosen), out_f has 8 floats; then per child (max_children): move, traversals,
// policy_prob, child score (-1 unexpanded), next_prob. Returns the number of children. Synchronises.
__global__ void puct_snapshot_kernel(const __grid_constant__ PuctEngine E, int g, int max_children, int* oi, float* of,
                                     int* moves, int* trav, float* prob, float* score, float* next_prob) {
    GameView V = gzp_view(&E, g);
    const GameCtl& c = *V.c;
    oi[0] = c.cmd_done; oi[1] = c.phase; oi[2] = c.error; oi[3] = c.choice; oi[4] = c.n_nodes; oi[5] = c.n_children;
    oi[6] = 0; oi[7] = 0; oi[8] = 0;
    for (int r = 0; r < GZP_MAX_ROLES; r++) oi[9 + r] = -1;
    of[0] = 0.0f;
    for (int r = 0; r < GZP_MAX_ROLES; r++) of[1 + r] = -1.0f;
    if (c.root < 0 || c.n_nodes <= 0) return;
    const int n = c.root;
    const int nc = NODE(n_nchild, n), c0 = NODE(n_child0, n), lead = NODE(n_lead, n);
    oi[6] = (int)NODE(n_visits, n); oi[7] = lead; oi[8] = nc;
    for (int r = 0; r < E.R && r < GZP_MAX_ROLES; r++) oi[9 + r] = NOTHER(n, r);   // non-lead roles' only action
    of[0] = NSCORE(n, lead);
    if (c.choice >= 0 && c.choice < nc) {        // the chosen child's node: its score for EVERY role (player.cpp:105-119)
        const int to = CHILD(c_to, c0 + c.choice);
        for (int r = 0; r < E.R && r < GZP_MAX_ROLES; r++) of[1 + r] = to >= 0 ? NSCORE(to, r) : -1.0f;
    }
    for (int i = 0; i < nc && i < max_children; i++) {
        moves[i] = CHILD(c_move, c0 + i);
        trav[i] = (int)CHILD(c_trav, c0 + i);
        prob[i] = CHILD(c_prob, c0 + i);
        const int to = CHILD(c_to, c0 + i);
        score[i] = to >= 0 ? NSCORE(to, lead) : -1.0f;
        next_prob[i] = CHILD(c_next, c0 + i);
    }
}

extern "C" int gz_puct_snapshot(gz_puct* h, int g, int max_children, int* out_i, float* out_f, int* moves, int* trav,
                                float* prob, float* score, float* next_prob, void* stream) {
    if (!h || g < 0 || g >= h->E.G || max_children < 0) { gz_set_error("bad argument"); return GZ_ERR_ARG; }
    GZ_CUDA(cudaSetDevice(h->device));
    { int jrc = join_side(h, (cudaStream_t)stream); if (jrc != GZ_OK) return jrc; }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t bytes = 16 * 4 + 8 * 4 + (size_t)max_children * 5 * 4 + 64;
    void* d = nullptr;
    GZ_CUDA(cudaMalloc(&d, bytes));
    int* di = (int*)d;
    float* df = (float*)(di + 16);
    int* dm = (int*)(df + 8);
    int* dt = dm + max_children;
    float* dp = (float*)(dt + max_children);
    float* ds = dp + max_children;
    float* dn = ds + max_children;
    puct_snapshot_kernel<<<1, 1, 0, s>>>(h->E, g, max_children, di, df, dm, dt, dp, ds, dn);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_i, di, 16 * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_f, df, 8 * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && max_children) {
        cudaMemcpyAsync(moves, dm, max_children * 4, cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(trav, dt, max_children * 4, cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(prob, dp, max_children * 4, cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(score, ds, max_children * 4, cudaMemcpyDeviceToHost, s);
        e = cudaMemcpyAsync(next_prob, dn, max_children * 4, cudaMemcpyDeviceToHost, s);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d);
    if (e != cudaSuccess) { gz_set_error("snapshot failed: %s", cudaGetErrorString(e)); return GZ_ERR_CUDA; }
    return out_i[8];
}

// Player::treeDebugInfo / PuctNode::debug (player.cpp:137-153, node.cpp:376-441): the root's children ordered by
// traversals, then by their node's score for the root's lead role; for each of the first max_count that is expanded:
// score, lead role, move, and the principal variation below it — at every node with at least 100 visits the child
// with the most traversals (ties: larger prior), up to `depth` plies — as (lead role, move) pairs.
// out_i per line: [lead, move, n_var, var pairs (2 * depth)]; out_f per line: score. Returns lines written.
__global__ void puct_tree_debug_kernel(const __grid_constant__ PuctEngine E, int g, int max_count, int depth,
                                       int* out_n, int* out_i, float* out_f) {
    GameView V = gzp_view(&E, g);
    const GameCtl& c = *V.c;
    *out_n = 0;
    if (c.root < 0 || c.n_nodes <= 0) return;
    const int n = c.root;
    const int nc = NODE(n_nchild, n), c0 = NODE(n_child0, n), lead = NODE(n_lead, n);
    const int stride = 3 + 2 * depth;
    int lines = 0;
    // selection of the next child in sortedChildrenDebug order: strictly after (prev_trav, prev_score, prev_index)
    long long prev_trav = 1ll << 40;
    float prev_score = 3.0e38f;
    int prev_idx = -1;
    const int limit = max_count < nc ? max_count : nc;
    for (int rank = 0; rank < limit; rank++) {
        int best = -1;
        long long bt = -1;
        float bs = -3.0e38f;
        for (int i = 0; i < nc; i++) {
            const long long t = (long long)CHILD(c_trav, c0 + i);
            const int to = CHILD(c_to, c0 + i);
            const float sc = to >= 0 ? NSCORE(to, lead) : -1.0f;
            // comes after the previous pick?
            const bool after = t < prev_trav || (t == prev_trav && (sc < prev_score || (sc == prev_score && i > prev_idx)));
            if (!after) continue;
            if (best < 0 || t > bt || (t == bt && sc > bs)) { best = i; bt = t; bs = sc; }
        }
        if (best < 0) break;
        prev_trav = bt; prev_score = bs; prev_idx = best;
        const int to0 = CHILD(c_to, c0 + best);
        if (to0 < 0) continue;                      // PuctNode::debug leaves an unexpanded child's line empty
        int* li = out_i + lines * stride;
        li[0] = lead;
        li[1] = CHILD(c_move, c0 + best);
        out_f[lines] = NSCORE(to0, lead);
        int nv = 0, cur = to0;
        for (int d = 0; d < depth; d++) {
            if (cur < 0 || NODE(n_nchild, cur) == 0 || NODE(n_visits, cur) < 100u) break;
            const int kc = NODE(n_nchild, cur), k0 = NODE(n_child0, cur);
            int top = 0;
            for (int i = 1; i < kc; i++) {
                const uint32_t ti = CHILD(c_trav, k0 + i), tt = CHILD(c_trav, k0 + top);
                if (ti > tt || (ti == tt && CHILD(c_prob, k0 + i) > CHILD(c_prob, k0 + top))) top = i;
            }
            li[3 + 2 * nv] = NODE(n_lead, cur);
            li[4 + 2 * nv] = CHILD(c_move, k0 + top);
            nv++;
            cur = CHILD(c_to, k0 + top);
        }
        li[2] = nv;
        lines++;
    }
    *out_n = lines;
}

extern "C" int gz_puct_tree_debug(gz_puct* h, int g, int max_count, int depth, int* h_out_i, float* h_out_f, void* stream) {
    if (!h || g < 0 || g >= h->E.G || max_count < 0 || depth < 0 || !h_out_i || !h_out_f) { gz_set_error("bad argument"); return GZ_ERR_ARG; }
    if (max_count == 0) return 0;
    GZ_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    { int jrc = join_side(h, s); if (jrc != GZ_OK) return jrc; }
    const int stride = 3 + 2 * depth;
    const size_t bytes = 16 + (size_t)max_count * stride * 4 + (size_t)max_count * 4;
    void* d = nullptr;
    GZ_CUDA(cudaMalloc(&d, bytes));
    int* dn = (int*)d;
    int* di = dn + 4;
    float* df = (float*)(di + (size_t)max_count * stride);
    puct_tree_debug_kernel<<<1, 1, 0, s>>>(h->E, g, max_count, depth, dn, di, df);
    int lines = 0;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(&lines, dn, 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_out_i, di, (size_t)max_count * stride * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_out_f, df, (size_t)max_count * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d);
    if (e != cudaSuccess) { gz_set_error("tree debug failed: %s", cudaGetErrorString(e)); return GZ_ERR_CUDA; }
    return lines;
}

extern "C" int gz_synth_nn(const void* d_planes, int plane_dtype, int64_t n, int total, uint64_t seed, int n_roles,
                           const int* h_role_sizes, int n_rewards, float* d_policy, float* d_value, void* stream) {
    if (!d_planes || !d_policy || !d_value || n_roles < 1 || n_roles > GZP_MAX_ROLES) { gz_set_error("bad argument"); return GZ_ERR_ARG; }
    if (n <= 0) return GZ_OK;
    static thread_local int* d_rso = nullptr;
    static thread_local int cached[2 * GZP_MAX_ROLES + 1] = {0};
    int rso[2 * GZP_MAX_ROLES];
    int off = 0;
    for (int r = 0; r < n_roles; r++) { rso[2 * r] = h_role_sizes[r]; rso[2 * r + 1] = off; off += h_role_sizes[r]; }
    if (!d_rso) GZ_CUDA(cudaMalloc(&d_rso, sizeof(rso)));
    if (memcmp(cached, rso, sizeof(int) * 2 * n_roles) != 0) {
        GZ_CUDA(cudaMemcpy(d_rso, rso, sizeof(int) * 2 * n_roles, cudaMemcpyHostToDevice));
        memcpy(cached, rso, sizeof(int) * 2 * n_roles);
    }
    const int wpb = 4;
    const int blocks = (int)((n + wpb - 1) / wpb);
    if (plane_dtype == GZ_DTYPE_BF16)
        synth_nn_kernel<__nv_bfloat16><<<blocks, wpb * 32, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)d_planes, n, total, seed, n_roles, d_rso, off, n_rewards, d_policy, d_value);
    else
        synth_nn_kernel<float><<<blocks, wpb * 32, 0, (cudaStream_t)stream>>>(
            (const float*)d_planes, n, total, seed, n_roles, d_rso, off, n_rewards, d_policy, d_value);
    GZ_CUDA(cudaGetLastError());
    return GZ_OK;
}
