/* gz_b200.h — C-ABI of the B200-native self-play hot path (libgz_b200.so).
 *
 * Plain pointers and sizes only; no torch / CUDA types in the signatures (streams cross as void*,
 * i.e. a cudaStream_t; 0 = the legacy default stream). Pointers named d_* are DEVICE pointers,
 * h_* are HOST pointers. All functions return 0 on success, a negative gz_status otherwise, and
 * gz_last_error() describes the failure. There is no CPU fallback: without a CUDA device every
 * compute entry point fails with GZ_ERR_CUDA.
 *
 * What each entry point replaces in the reference (the calls a maintainer would re-bind):
 *
 *   gz_propnet_create            ggplib's state-machine construction for a game, consumed through
 *                                sm_ptr by Supervisor/Player (src/cpp/pyobjects/supervisor_impl.cpp:215-237,
 *                                src/cpp/supervisor.cpp:20-31 `sm->dupe()`)
 *   gz_sm_eval                   StateMachineInterface::updateBases + getLegalState + isTerminal +
 *                                getGoalValue            (src/cpp/puct/node.cpp:157-217), batched
 *   gz_sm_transition             updateBases + nextState + updateBases + legal/terminal/goal
 *                                (src/cpp/puct/evaluator.cpp:216-239 expandChild -> node.cpp:153-221)
 *   gz_sm_playout_step           the random playout loop of GdlBasesTransformer.test
 *                                (src/cpp/pyobjects/gdltransformer_impl.cpp:50-124), one ply per call
 *   gz_sm_transition_host        same as gz_sm_transition with host buffers (copies inside)
 *   gz_transformer_* / gz_encode GdlBasesTransformer and toChannels (src/cpp/gdltransformer.cpp:14-52)
 *   gz_puct_*                    Supervisor / Player: create, advance (device tensors), host_poll_submit / _wait
 *                                (numpy poll contract), fetch_samples, command / snapshot (src/cpp/supervisor.cpp,
 *                                src/cpp/player.cpp and the files under src/cpp/pyobjects)
 *   gz_permute_state_bits        Translator.translate_basestate, batched (src/ggpzero/util/symmetry.py:153-204)
 */
#ifndef GZ_B200_H
#define GZ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    GZ_OK = 0,
    GZ_ERR_ARG = -1,
    GZ_ERR_CUDA = -2,
    GZ_ERR_FORMAT = -3,
    GZ_ERR_LIMIT = -4
} gz_status;

typedef struct gz_propnet gz_propnet;

/* goals_mode */
#define GZ_GOALS_ALWAYS 0        /* goal values computed for every state */
#define GZ_GOALS_TERMINAL_ONLY 1 /* goal cone only evaluated for tiles holding a terminal state;
                                    non-terminal states report 255 */

const char* gz_last_error(void);
int gz_version(void);

/* Builds the device-resident program of one game from a device blob (galvanise_zero_b200/devprog.py,
 * layout in csrc/gz_devformat.h) on CUDA device `device`. */
int gz_propnet_create(const int32_t* h_dev_blob, int64_t n_words, int device, gz_propnet** out);
void gz_propnet_destroy(gz_propnet* pn);

/* info selectors for gz_propnet_info */
enum {
    GZ_INFO_N_BASES = 0, GZ_INFO_N_ROLES = 1, GZ_INFO_STATE_WORDS = 2, GZ_INFO_LEGAL_WORDS = 3,
    GZ_INFO_VAL_WORDS = 4, GZ_INFO_WARPS_PER_CTA = 5, GZ_INFO_SMEM_BYTES = 6, GZ_INFO_GRID = 7,
    GZ_INFO_STAGED_MASK = 8, GZ_INFO_N_INPUTS = 9, GZ_INFO_SPECIALISED = 10, GZ_INFO_ROLE_SIZE0 = 16 /* +r */,
    GZ_INFO_LEGAL_WORD_OFF0 = 24 /* +r */
};
int64_t gz_propnet_info(const gz_propnet* pn, int what);
/* Games with a generated, game-specialised kernel (csrc/gz_spec_<game>.cu; GZ_INFO_SPECIALISED == 1) use it
 * for batches of at least `min_batch` states (default 2^18); min_batch < 0 disables it, 0 forces it. The
 * table-driven kernel and the specialised one return identical results. */
int gz_propnet_set_spec_min_batch(gz_propnet* pn, int64_t min_batch);
/* Wide-format games (chess_15d, hexLG13) evaluate batches of at least `min_batch` states (default 1024) with the
 * group kernel (1024 games per CTA / cluster, one tile per lane); < 0 disables it, 0 forces it; other games only
 * when asked. Identical results. */
int gz_propnet_set_group_min_batch(gz_propnet* pn, int64_t min_batch);
/* GZ_GOALS_TERMINAL_ONLY on a game with a deep goal cone (reversi's piece counting): batches of at least `min_batch`
 * states (default 4096; < 0: never) skip the cone in the main pass and evaluate it in a second pass over the
 * compacted terminal states only — the reference reads goal values of terminal states only
 * (src/cpp/puct/node.cpp:153-221). Identical results. */
int gz_propnet_set_deferred_goals_min_batch(gz_propnet* pn, int64_t min_batch);

/* Layouts (n = number of states, R = roles):
 *   states  uint32 [n][state_words]   packed bases, LSB first, rows padded to 16 B
 *   legal   uint32 [n][legal_words]   per role ceil(A_r/32) words at legal_word_off[r], padded to 16 B
 *   moves   int16  [n][R]             one action index per role
 *   terminal uint8 [n]     goals uint8 [n][R] (255 = no goal proposition true / not evaluated)
 */
int gz_sm_eval(gz_propnet* pn, int64_t n, const uint32_t* d_states, uint32_t* d_legal,
               uint8_t* d_terminal, uint8_t* d_goals, int goals_mode, void* stream);

int gz_sm_transition(gz_propnet* pn, int64_t n, const uint32_t* d_states, const int16_t* d_moves,
                     uint32_t* d_next, uint32_t* d_legal, uint8_t* d_terminal, uint8_t* d_goals,
                     int goals_mode, void* stream);

/* One random-playout ply for slots [slot0, slot0+n), in place. Terminal slots restart from the
 * initial state; the others draw one uniformly random legal action per role with
 * gz_rand32(seed, slot, step, role) (include/gz_propnet_format.h) and advance.
 * d_moves_out (optional, may be NULL) receives the chosen actions (-1 on restart);
 * d_transitions (optional) is incremented by the number of real transitions. */
int gz_sm_playout_step(gz_propnet* pn, int64_t slot0, int64_t n, uint64_t seed, uint64_t step,
                       uint32_t* d_states, uint32_t* d_legal, uint8_t* d_terminal, uint8_t* d_goals,
                       int16_t* d_moves_out, unsigned long long* d_transitions, int goals_mode,
                       void* stream);

/* Host-buffer variant of gz_sm_transition: H2D copies, kernel, D2H copies and a stream
 * synchronise happen inside the call (pinned staging owned by `pn`). */
int gz_sm_transition_host(gz_propnet* pn, int64_t n, const uint32_t* h_states, const int16_t* h_moves,
                          uint32_t* h_next, uint32_t* h_legal, uint8_t* h_terminal, uint8_t* h_goals,
                          int goals_mode);


/* ------------------------------------------------------------------------------------------------
 * Base-state -> input-plane encoder. Replaces GdlBasesTransformer (src/cpp/gdltransformer.h:69-149,
 * exported as ggpzero_interface.GdlBasesTransformer, src/cpp/pyobjects/gdltransformer_impl.cpp:186-219):
 *   gz_transformer_create            the constructor "iiiiiO!" (channel_size, channels_per_state,
 *                                    num_control_channels, num_prev_states, num_rewards, [policy sizes])
 *   gz_transformer_add_board_base    .add_board_base(base_idx, buf_offset)     (gdltransformer_impl.cpp:27-36)
 *   gz_transformer_add_control_base  .add_control_base(base_idx, channel, value) (gdltransformer_impl.cpp:38-48)
 *   gz_transformer_finalize          (new) freezes the tables and uploads them to `device`
 *   gz_encode                        GdlBasesTransformer::toChannels (src/cpp/gdltransformer.cpp:22-52), batched
 */
typedef struct gz_transformer gz_transformer;
#define GZ_DTYPE_F32 0
#define GZ_DTYPE_BF16 1

int gz_transformer_create(int channel_size, int channels_per_state, int num_control_channels,
                          int num_prev_states, int num_rewards, const int* policy_sizes, int n_policies,
                          gz_transformer** out);
void gz_transformer_destroy(gz_transformer* t);
int gz_transformer_add_board_base(gz_transformer* t, int base_idx, int buf_offset);
int gz_transformer_add_control_base(gz_transformer* t, int base_idx, int channel_id, float value);
int gz_transformer_finalize(gz_transformer* t, int n_bases, int device);
int gz_transformer_total_size(const gz_transformer* t);

/* d_states uint32 [n][state_words]; d_prev uint32 [n][max_prev][state_words] (nearest ancestor first,
 * may be NULL); d_prev_count uint8 [n] = how many of them exist (NULL = all max_prev);
 * d_out [n][total_size] of out_dtype, channels-first. */
int gz_encode(gz_transformer* t, int64_t n, int state_words, const uint32_t* d_states,
              const uint32_t* d_prev, const uint8_t* d_prev_count, int max_prev, void* d_out,
              int out_dtype, void* stream);
/* same, with the output layout selectable: c_pad == 0 -> NCHW rows; c_pad > 0 -> [n][H][W][c_pad] */
int gz_encode_layout(gz_transformer* t, int64_t n, int state_words, const uint32_t* d_states,
                     const uint32_t* d_prev, const uint8_t* d_prev_count, int max_prev, void* d_out,
                     int out_dtype, int c_pad, void* stream);

/* Batches of at least `min_batch` states whose plane rows are a multiple of 8 elements are encoded by the
 * column-owned kernel (csrc/gz_encoder_col.cuh; default 16384, negative: never) — same results, a tuning switch. */
int gz_encode_set_column_kernel_min_batch(int64_t min_batch);
/* measurement switch of that kernel: states per tile (0: default) */
int gz_encode_set_column_kernel_tile(int tile_states);
/* rows that are NOT a multiple of 8 elements (hexLG13, breakthroughSmall) through the column kernel as well, in units
 * of 8 / gcd(8, row) states: -1 (default) where a tile of states fits 48 KB of shared memory (breakthroughSmall),
 * 1 always (hexLG13: opt-in shared memory; bit-exact on a B200, not measured), 0 never */
int gz_encode_set_column_kernel_ragged(int on);

/* Board symmetry of packed states, batched (src/ggpzero/util/symmetry.py:153-204 translate_basestate): bit j of
 * every output state = bit d_src[j] of the input state (d_src: int32 [n_bases], a permutation; padding bits
 * come out zero). d_out must not alias d_states. */
int gz_permute_state_bits(int64_t n, int state_words, int n_bases, const uint32_t* d_states, const int* d_src,
                          uint32_t* d_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Batched PUCT self-play engine. Replaces the reference's Supervisor / SelfPlayManager / SelfPlay /
 * NetworkScheduler / PuctEvaluator stack for `n_games` concurrent games:
 *   gz_puct_create          ggpzero_interface.Supervisor(sm_ptr, transformer, batch_size, identifier)
 *                           + .start_self_play(num_workers, conf_dict)
 *                           (src/cpp/pyobjects/supervisor_impl.cpp:14-32,215-237); mode 1 = Player(...)
 *                           (src/cpp/pyobjects/player_impl.cpp:183-201)
 *   gz_puct_advance         Supervisor.poll(n_predictions, [policy_0.., value]) -> planes
 *                           (src/cpp/pyobjects/common.cpp:160-218, src/cpp/supervisor.cpp:101-168), with
 *                           device tensors on both sides (no host round trip)
 *   gz_puct_fetch_samples   Supervisor.fetch_samples()           (supervisor_impl.cpp:75-128)
 *   gz_puct_clear_unique_states  .clear_unique_states()          (supervisor_impl.cpp:138-144)
 *   gz_puct_command / gz_puct_snapshot   Player.player_reset / player_move / player_apply_move /
 *                           player_get_move / player_tree_debug (src/cpp/pyobjects/player_impl.cpp:8-129)
 *   gz_puct_totals          the counters of SelfPlayManager::reportAndResetStats
 *                           (src/cpp/selfplaymanager.cpp:161-200) plus throughput counters
 * Config arrays are flat doubles in the field order of oracle/gzref.py SP_FIELDS / PUCT_FIELDS, i.e.
 * the fields of SelfPlayConfig / PUCTEvaluatorConfig (src/ggpzero/defs/confs.py:9-123).
 */
typedef struct gz_puct gz_puct;
#define GZ_PUCT_MODE_SELFPLAY 0
#define GZ_PUCT_MODE_PLAYER 1
#define GZ_PUCT_CMD_RESET 1
#define GZ_PUCT_CMD_APPLY 2
#define GZ_PUCT_CMD_SEARCH 3
#define GZ_PUCT_CMD_BALANCE 4   /* arg = how many of the root's first moves (PuctEvaluator::balanceFirstMoves, evaluator.cpp:282-339) */

int gz_puct_create(gz_propnet* pn, gz_transformer* tr, int n_games, int mode, int node_cap, int child_cap,
                   int max_pending_samples, int ready_cap, const double* sp_conf, const double* puct_conf,
                   const double* run_to_end_puct_conf, uint64_t seed, gz_puct** out);
void gz_puct_destroy(gz_puct* h);
int gz_puct_set_config(gz_puct* h, const double* sp_conf, const double* puct_conf, const double* run_to_end_puct_conf);
/* what: 0 games, 1 plane row size, 2 policy row size (sum of role sizes), 3 value row size,
 *       4 sample record words, 5 arena bytes, 6 state words, 7 max actions, 8 ready capacity,
 *       9 kernels of this library launched per engine step */
int64_t gz_puct_info(const gz_puct* h, int what);
int gz_puct_advance(gz_puct* h, const float* d_policy, const float* d_value, int nn_fresh, int sub_iters,
                    void* d_planes, int plane_dtype, void* stream);
/* gz_puct_advance forks part of a step onto an internal side stream (re-rooting of the games that moved; for
 * games with a deep goal cone, the goal values and nodes of terminal states) and joins it before returning
 * control of `stream`. The _async variant leaves that work running so that the caller's network forward
 * overlaps it; gz_puct_join (or the next advance, which joins first) must follow on the same stream before
 * the engine is read or stepped again. */
int gz_puct_advance_async(gz_puct* h, const float* d_policy, const float* d_value, int nn_fresh, int sub_iters,
                          void* d_planes, int plane_dtype, void* stream);
int gz_puct_join(gz_puct* h, void* stream);
/* Measurement aid: one engine step with CUDA events between its launches; waits for it and returns the
 * device time of each launch in ms: [puct_pre_kernel, puct_reroot_kernel, pn_kernel, puct_post_kernel,
 * encode_kernel]. */
int gz_puct_advance_timed(gz_puct* h, const float* d_policy, const float* d_value, int nn_fresh,
                          void* d_planes, int plane_dtype, void* stream, float* h_ms5);
/* planes written by gz_puct_advance: c_pad == 0 -> [G][C][H][W] (reference layout, default);
 * c_pad > 0 (multiple of 8) -> channels-last [G][H][W][c_pad], zero planes beyond C */
int gz_puct_set_plane_layout(gz_puct* h, int c_pad);
const uint8_t* gz_puct_need_eval_ptr(const gz_puct* h);
int gz_puct_copy_need_eval(gz_puct* h, uint8_t* d_out, void* stream);
/* Host-buffer poll — Supervisor.poll(n_predictions, [policy_0.., value]) with HOST arrays on both sides
 * (src/cpp/pyobjects/common.cpp:160-218): h_arrays = R policy arrays [n_pred][A_r] + one value array
 * [n_pred][rewards], float32, rows in the order handed out by the previous wait (n_pred must equal it;
 * 0 on the first call). submit() enqueues copy-in, the engine step, compaction and copy-out on `stream`
 * and returns; wait() blocks until they finished and returns the number of rows now waiting for the
 * network, *h_planes = [n_req][total_size] float32 in pinned memory owned by the engine (valid until
 * the next submit). Two engines polled alternately overlap one's step with the network call on the
 * other's batch (the reference's two managers per worker, src/cpp/supervisor.cpp:79-99). */
int gz_puct_host_poll_submit(gz_puct* h, int n_pred, const float* const* h_arrays, int n_arrays, void* stream);
int gz_puct_host_poll_wait(gz_puct* h, const float** h_planes);
/* h_out26: evals, moves, games, samples, playouts, errors, halted, first_error, max_nodes, max_children, dupes,
 * no_samples, fp_resign0, fp_resign1, early_run_to_ends, resigns, aborts, waiting_eval, ready, select_nodes,
 * select_children, new_children, depth_sum (sum over slots of the current root's game depth), dropped_samples
 * (finished samples lost because the ready buffer was full: fetch more often), error_restarts (games abandoned at a
 * per-game capacity limit; their slot started a new game), transpositions (expansions that attached an existing node:
 * PuctConfig.lookup_transpositions, stats.num_transpositions_attached, src/cpp/puct/evaluator.cpp:224-226). */
int gz_puct_totals(gz_puct* h, double* h_out26, void* stream);
/* Measurement aid: the next game of every slot starts from d_states[g] (device, [G][state_words], non-terminal
 * positions) at game depth 0 instead of the initial state — what handing SelfPlay another `initial_state`
 * (src/cpp/selfplaymanager.cpp:140, selfplay.cpp:313) does in the reference. Later games start normally. */
int gz_puct_set_start_states(gz_puct* h, const uint32_t* d_states, void* stream);
/* Compacted network batches on the device (self-play engines). After a step, gz_puct_compact_requests lists the rows
 * that wait for the network (ascending game order), keeps at most `cap` of them and gathers their planes (rows of
 * row_bytes bytes as gz_puct_advance wrote them) into d_planes_c; h_count_pinned (optional, pinned
 * host memory) receives the uncapped number of waiting rows with an asynchronous copy. gz_puct_advance_compact is the
 * next step: d_policy_c [n][policy_stride] / d_value_c [n][rewards] are the network's outputs for those rows in that
 * order. Requests beyond the cap stay pending (their games wait one more step), so the batch may be sized from a
 * stale count without synchronising. This is what the reference's poll does with host arrays
 * (src/cpp/supervisor.cpp:101-168: only the waiting games' planes are handed out). */
int gz_puct_compact_requests(gz_puct* h, int cap, const void* d_planes, int64_t row_bytes, void* d_planes_c,
                             int* h_count_pinned, void* stream);
int gz_puct_advance_compact(gz_puct* h, const float* d_policy_c, const float* d_value_c, int sub_iters, void* d_planes,
                            int plane_dtype, void* stream, int defer_join);
/* Measurement aid: d_buf = device int64 [G][4] (or null to stop): per game the cycles its warp spent in the last
 * puct_pre_kernel launch, entry*256+exit phase, playouts finished, moves*65536 + path_len*256 + request. */
int gz_puct_set_debug_cycles(gz_puct* h, long long* d_buf);
/* PuctNode::debug for game g's root (src/cpp/puct/node.cpp:376-441): per line h_out_i[(3 + 2*depth) * i] = lead role,
 * move, n pairs, then the (lead role, move) pairs of the variation; h_out_f[i] = score. Returns the lines written. */
int gz_puct_tree_debug(gz_puct* h, int g, int max_count, int depth, int* h_out_i, float* h_out_f, void* stream);
/* Per engine step a game runs until it needs the batch pipeline (a propnet pass / the network), but at most
 * `max_iters` controller iterations and `max_work` tree levels descended (playouts that end in terminal or
 * finalised nodes need no pipeline pass; the slowest game sets the duration of the step). Defaults 16 / 8.
 * No reference counterpart: the reference runs one coroutine per game (src/cpp/scheduler.cpp). */
int gz_puct_set_step_limits(gz_puct* h, int max_iters, int max_work);
int64_t gz_puct_fetch_samples(gz_puct* h, int32_t* h_out, int64_t cap, void* stream);
int gz_puct_clear_unique_states(gz_puct* h, void* stream);
int gz_puct_command(gz_puct* h, int game, int cmd, int arg, int evals, const uint32_t* h_state, void* stream);
int gz_puct_snapshot(gz_puct* h, int game, int max_children, int* out_i16, float* out_f8, int* moves, int* trav,
                     float* prob, float* score, float* next_prob, void* stream);

/* Player mode: one 16-byte status copy of game `game` — [cmd_done, phase, error, waiting for the network] — and
 * the host's clock talking to a running search (mode 1: stop at the next playout boundary — the hard end_time, or
 * think_time x multiplier; mode 2: stop as soon as the search has converged — think_time elapsed;
 * src/cpp/puct/evaluator.cpp:794,820-826). */
int gz_puct_status(gz_puct* h, int game, int* h_out4, void* stream);
int gz_puct_request_stop(gz_puct* h, int game, int mode, void* stream);

/* ---- the reference's Supervisor and Player objects, native host side (csrc/gz_host.cu) -----------------------
 * What src/cpp/pyobjects/supervisor_impl.cpp and player_impl.cpp bind, method for method; the extension module
 * galvanise_zero_b200/csrc/ggpzero_interface.cpp and the ctypes host galvanise_zero_b200/interface.py are both thin
 * layers over these calls.
 *
 *   gz_supervisor_create            Supervisor(sm_ptr, transformer, batch_size, identifier)  supervisor_impl.cpp:215-237
 *   gz_supervisor_start_self_play   .start_self_play(num_workers, conf_dict)                 supervisor_impl.cpp:14-32
 *                                   (configuration as the flat arrays of gz_puct_create; num_workers <= 0: inline,
 *                                   one manager; > 0: worker mode, two managers, supervisor.cpp:68-99)
 *   gz_supervisor_poll              .poll(n_predictions, [policy_0.., value])                common.cpp:160-218
 *                                   returns rows handed out (> 0, *h_planes pinned [rows][total] float32), 0 = None
 *   gz_supervisor_fetch_samples     .fetch_samples()  (records; the binding turns them into dicts)  :75-128
 *   gz_supervisor_clear_unique_states                                                        :138-144
 *   gz_player_create                Player(sm_ptr, transformer, puct_conf_dict)              player_impl.cpp:183-201
 *   gz_player_reset / _apply_move / _move / _get_move / _update_config / _balance_moves / _tree_debug / _poll
 *                                   the player_* methods                                      player_impl.cpp:8-129
 */
typedef struct gz_supervisor gz_supervisor;
int gz_supervisor_create(gz_propnet* pn, gz_transformer* tr, int batch_size, const char* identifier, int device,
                         uint64_t seed, int node_cap, int child_cap, gz_supervisor** out);
void gz_supervisor_destroy(gz_supervisor* s);
int gz_supervisor_start_self_play(gz_supervisor* s, int num_workers, const double* sp_conf, const double* puct_conf,
                                  const double* run_to_end_puct_conf);
int64_t gz_supervisor_info(const gz_supervisor* s, int what); /* 0 managers, 1 batch size, 2 plane row, 3 record words, 4 roles */
gz_puct* gz_supervisor_engine(gz_supervisor* s, int manager);
void* gz_supervisor_stream(gz_supervisor* s, int manager);
int gz_supervisor_poll(gz_supervisor* s, int n_pred, const float* const* h_arrays, int n_arrays, const float** h_planes);
int64_t gz_supervisor_fetch_samples(gz_supervisor* s, int32_t* h_out, int64_t cap, int32_t* h_manager);
int gz_supervisor_clear_unique_states(gz_supervisor* s);
int gz_supervisor_totals(gz_supervisor* s, double* h_out26);

typedef struct gz_player gz_player;
int gz_player_create(gz_propnet* pn, gz_transformer* tr, const double* puct_conf, int device, uint64_t seed, int node_cap,
                     int child_cap, gz_player** out);
void gz_player_destroy(gz_player* p);
gz_puct* gz_player_engine(gz_player* p);
int64_t gz_player_info(const gz_player* p, int what);         /* 0 plane row, 1 roles, 2 max actions, 3 state words, 4 most rows per poll */
int gz_player_reset(gz_player* p, int game_depth);
int gz_player_update_config(gz_player* p, double think_time, int converged_visits, int verbose);
int gz_player_apply_move(gz_player* p, const int* joint_move, int n_roles);
int gz_player_move(gz_player* p, const uint32_t* h_state, int evaluations, double end_time);
int gz_player_get_move(gz_player* p, int role, int* move, float* score, int* nodes);
int gz_player_balance_moves(gz_player* p, int max_count);
int gz_player_tree_debug(gz_player* p, int max_count, float* scores, int* lead, int* moves);
/* ... with the variation lines of PuctNode::debug (src/cpp/puct/node.cpp:400-441): var_len[i] (lead role, move) pairs
 * at var_pairs + 20 * i (at most 10 plies). */
int gz_player_tree_debug_lines(gz_player* p, int max_count, float* scores, int* lead, int* moves, int* var_len,
                               int* var_pairs);
int gz_player_poll(gz_player* p, int n_pred, const float* const* h_arrays, int n_arrays, const float** h_planes);
int gz_player_last_search(gz_player* p, int* out_i16, float* out_f8, int* moves, int* trav, float* prob, float* score,
                          float* next_prob, int cap);

/* ------------------------------------------------------------------------------------------------
 * Sample records (what gz_puct_fetch_samples / gz_supervisor_fetch_samples return, int32 words, `record words`
 * of gz_puct_info per record) = the fields of datadesc.Sample (src/ggpzero/defs/datadesc.py:8-54) as
 * Supervisor.fetch_samples hands them out (src/cpp/pyobjects/supervisor_impl.cpp:75-118):
 * a header of GZ_SR_HEADER words, the state row, the previous-state row (valid when GZ_SR_N_PREV is 1), then
 * GZ_SR_N_CHILDREN (action index of the lead role, probability as float bits) pairs. */
#define GZ_SR_DEPTH 1
#define GZ_SR_GAME_LENGTH 2
#define GZ_SR_SLOT 3
#define GZ_SR_MATCH 4
#define GZ_SR_HAS_RESIGNED 5
#define GZ_SR_RESIGN_FP 6
#define GZ_SR_START_DEPTH 7
#define GZ_SR_VISITS 8
#define GZ_SR_LEAD 9          /* role whose policy the pairs are; -1: no lead role */
#define GZ_SR_N_PREV 10
#define GZ_SR_N_CHILDREN 11
#define GZ_SR_FINAL_SCORE 12  /* float[4] */
#define GZ_SR_PUCT_SCORE 16   /* float[4] */
#define GZ_SR_OTHER_MOVES 20  /* int[4]: the single move of every role that is not the lead role */
#define GZ_SR_HEADER 24

/* Sample records -> training arrays, batched on the device: what the reference's trainer does per sample in
 * DataCache.sync (src/ggpzero/nn/datacache.py:700-760) with the transformer (src/ggpzero/nn/bases.py):
 * final scores clamped to [0, 1] (datacache.py:716-718), state_to_channels(state, prev_states) (bases.py:242-287)
 * -> d_planes [n][total_size] of plane_dtype, policy_to_array(policy, role) (bases.py:305-310) -> d_policy[r]
 * float [n][policy_sizes[r]] (the lead role's pairs; every other role its single move with probability 1,
 * src/cpp/selfplaymanager.cpp:95-108), value_to_array(final_score) (bases.py:312-314; with n_values == 3 the draw
 * head of GdlBasesTransformer_Draws, :323-335) -> d_value float [n][n_values]; d_bad (may be NULL) receives the
 * number of records that fail check_sample (bases.py:289-303: every probability in (-0.01, 1.01), the lead
 * role's sum in (0.99, 1.01), action indices inside the role's policy width; the draw head's own asserts).
 * Augmentation (DataCache.augment_data, datacache.py:597-671): with d_sym (int32 [n], one symmetry id per record,
 * 0 <= id < n_sym) the state rows go through d_base_src (int32 [n_sym][n_bases]: new bit j = old bit src[j], the
 * inverse of Translator.translate_basestate's map, src/ggpzero/util/symmetry.py:153-204) and every action index
 * through d_action_dest[r] (int32 [n_sym][policy_sizes[r]]: translate_action, symmetry.py:206-225); d_sym NULL:
 * records as they are. Output row i is made from record d_index[i] (int32 [n]; NULL: record i) of the n_records
 * records, so that one record can appear under several symmetries. d_scratch: gz_records_to_training_scratch_bytes(n, state_words) bytes. */
int64_t gz_records_to_training_scratch_bytes(int64_t n, int state_words);
int gz_records_to_training(gz_transformer* t, int64_t n, int64_t n_records, int record_words, int state_words,
                           int n_roles, const int32_t* d_records, const int32_t* d_index, const int32_t* d_sym,
                           int n_sym, const int32_t* d_base_src,
                           const int32_t* const* d_action_dest, void* d_planes, int plane_dtype,
                           float* const* d_policy, float* d_value, int n_values, int32_t* d_bad, void* d_scratch,
                           void* stream);

/* Deterministic synthetic policy/value network (include/gz_synth.h) on encoded planes — parity and
 * benchmarking utility; the production network is a PyTorch module. */
int gz_synth_nn(const void* d_planes, int plane_dtype, int64_t n, int total, uint64_t seed, int n_roles,
                const int* h_role_sizes, int n_rewards, float* d_policy, float* d_value, void* stream);

#ifdef __cplusplus
}
#endif
#endif
