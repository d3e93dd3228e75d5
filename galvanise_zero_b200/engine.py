"""Device PUCT self-play engine: Python handle over the gz_puct_* C-ABI (include/gz_b200.h).

One ``PuctEngine`` owns ``n_games`` game trees on one GPU. ``advance(policy, value)`` is the
device-tensor counterpart of the reference's ``Supervisor.poll(n, [policy_0.., value])``
(``src/ggpzero/util/cppinterface.py:78-129``): it feeds the network outputs for the previous batch
of leaves and returns the next batch of encoded planes — as CUDA tensors, without a host round trip.
"""
from __future__ import annotations

import ctypes

import attr
import numpy as np
import torch

from . import bases, cppinterface, lookup
from . import lib as _lib
from .defs import confs, datadesc
from .statemachine import DeviceStateMachine

PUCT_FIELDS = ['puct_constant', 'puct_constant_root', 'dirichlet_noise_pct', 'noise_policy_squash_pct',
               'noise_policy_squash_prob', 'choose', 'random_scale', 'temperature',
               'depth_temperature_start', 'depth_temperature_increment', 'depth_temperature_stop',
               'depth_temperature_max', 'fpu_prior_discount', 'fpu_prior_discount_root',
               'top_visits_best_guess_converge_ratio', 'think_time', 'converged_visits', 'batch_size',
               'use_legals_count_draw', 'backup_finalised', 'lookup_transpositions',
               'evaluation_multiplier_to_convergence']
SP_FIELDS = ['oscillate_sampling_pct', 'temperature_for_policy', 'evals_per_move',
             'resign0_score_probability', 'resign0_pct', 'resign1_score_probability', 'resign1_pct',
             'run_to_end_pct', 'run_to_end_evals', 'run_to_end_early_score',
             'run_to_end_minimum_game_depth', 'abort_max_length']
TOTAL_KEYS = ['evals', 'moves', 'games', 'samples', 'playouts', 'errors', 'halted', 'first_error',
              'max_nodes', 'max_children', 'dupes', 'no_samples', 'fp_resign0', 'fp_resign1',
              'early_run_to_ends', 'resigns', 'aborts', 'waiting_eval', 'ready',
              'select_nodes', 'select_children', 'new_children', 'depth_sum', 'dropped_samples',
              'error_restarts', 'transpositions']
# sample record header (csrc/gz_puct_core.h SR_*)
SR_DEPTH, SR_GAME_LENGTH, SR_SLOT, SR_MATCH, SR_HAS_RESIGNED, SR_RESIGN_FP, SR_START_DEPTH, SR_VISITS, \
    SR_LEAD, SR_N_PREV, SR_N_CHILDREN = range(1, 12)
SR_FINAL_SCORE, SR_PUCT_SCORE, SR_OTHER_MOVES, SR_HEADER = 12, 16, 20, 24


def _puct_array(d):
    vals = []
    for f in PUCT_FIELDS:
        v = d.get(f, attr.asdict(confs.PUCTEvaluatorConfig())[f])   # tolerate missing keys
        if f == 'choose':
            v = 1.0 if v == 'choose_temperature' else 0.0
        vals.append(float(v))
    return (ctypes.c_double * len(vals))(*vals)


def _sp_array(d):
    dflt = attr.asdict(confs.SelfPlayConfig())
    return (ctypes.c_double * len(SP_FIELDS))(*[float(d.get(f, dflt[f])) for f in SP_FIELDS])


def _declare(L):
    if getattr(L, '_gz_puct_declared', False):
        return
    vp, ci, i64, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64
    dp = ctypes.POINTER(ctypes.c_double)
    L.gz_puct_create.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, dp, dp, dp, u64, ctypes.POINTER(vp)]
    L.gz_puct_destroy.argtypes = [vp]
    L.gz_puct_set_config.argtypes = [vp, dp, dp, dp]
    L.gz_puct_info.argtypes = [vp, ci]
    L.gz_puct_info.restype = i64
    L.gz_puct_advance.argtypes = [vp, vp, vp, ci, ci, vp, ci, vp]
    L.gz_puct_advance_async.argtypes = [vp, vp, vp, ci, ci, vp, ci, vp]
    L.gz_puct_join.argtypes = [vp, vp]
    L.gz_puct_totals.argtypes = [vp, dp, vp]
    L.gz_puct_fetch_samples.argtypes = [vp, vp, i64, vp]
    L.gz_puct_fetch_samples.restype = i64
    L.gz_puct_clear_unique_states.argtypes = [vp, vp]
    L.gz_puct_set_start_states.argtypes = [vp, vp, vp]
    L.gz_puct_command.argtypes = [vp, ci, ci, ci, ci, vp, vp]
    L.gz_puct_snapshot.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp]
    L.gz_synth_nn.argtypes = [vp, ci, i64, ci, u64, ci, ctypes.POINTER(ci), ci, vp, vp, vp]
    L.gz_puct_compact_requests.argtypes = [vp, ci, vp, i64, vp, vp, vp]
    L.gz_puct_advance_compact.argtypes = [vp, vp, vp, ci, vp, ci, vp, ci]
    L._gz_puct_declared = True


class PuctEngine(object):
    MODE_SELFPLAY, MODE_PLAYER = 0, 1

    def __init__(self, game, sp_conf=None, n_games=1024, mode=0, device=0, seed=0x5EED0000,
                 node_cap=4096, child_cap=None, generation=None, sm=None, c_transformer=None,
                 max_pending_samples=128, ready_cap=1 << 16, plane_dtype=torch.bfloat16, plane_c_pad=0):
        self._L = _lib.lib()
        _declare(self._L)
        self.game = game
        self.game_info = lookup.by_name(game)
        self.generation = generation or datadesc.generation_for(game)
        self.transformer = bases.GdlBasesTransformer(self.game_info, self.generation)
        self.sm = sm or DeviceStateMachine(game, device=device, goals_mode=_lib.GOALS_TERMINAL_ONLY)
        self.c_transformer = c_transformer or cppinterface.create_c_transformer(self.transformer)
        self.c_transformer.finalize(self.transformer.num_bases, device)
        self.device = torch.device('cuda', device)
        sp_conf = sp_conf or confs.selfplay_config_template()
        self.sp_conf = sp_conf if isinstance(sp_conf, dict) else attr.asdict(sp_conf)
        max_actions = max(self.sm.role_sizes)
        if child_cap is None:
            child_cap = node_cap * max(8, min(max_actions, 64) // 2)
        h = ctypes.c_void_p()
        _lib.check(self._L.gz_puct_create(
            self.sm._h, self.c_transformer._h, n_games, mode, node_cap, child_cap, max_pending_samples,
            ready_cap, _sp_array(self.sp_conf), _puct_array(self.sp_conf['puct_config']),
            _puct_array(self.sp_conf['run_to_end_puct_config']), seed, ctypes.byref(h)))
        self._owns = True
        tr = self.transformer
        self._setup(h, n_games, mode, (tr.num_channels, tr.num_cols, tr.num_rows), tr.num_bases, list(self.sm.role_sizes),
                    plane_dtype, plane_c_pad)

    @classmethod
    def from_handle(cls, handle, device, n_bases, role_sizes, plane_shape, mode=0, plane_dtype=torch.float32, sm=None):
        """A view of an engine owned by somebody else (a native gz_supervisor / gz_player): same methods, no
        ownership. ``plane_shape`` = (channels, cols, rows)."""
        self = cls.__new__(cls)
        self._L = _lib.lib()
        _declare(self._L)
        self.game = getattr(sm, 'game', None)
        self.sm = sm
        self.transformer = None
        self.c_transformer = None
        self.device = torch.device('cuda', device)
        self._owns = False
        n_games = int(self._L.gz_puct_info(handle, 0))
        self._setup(handle, n_games, mode, tuple(plane_shape), n_bases, list(role_sizes), plane_dtype, 0)
        return self

    def _setup(self, h, n_games, mode, plane_shape, n_bases, role_sizes, plane_dtype, plane_c_pad):
        self._h = h if isinstance(h, ctypes.c_void_p) else ctypes.c_void_p(h)
        self.n_games = n_games
        self.mode = mode
        self.plane_dtype = plane_dtype
        self.n_bases = n_bases
        info = lambda w: int(self._L.gz_puct_info(self._h, w))  # noqa: E731
        self.total_size = info(1)
        self.policy_stride = info(2)
        self.n_rewards = info(3)
        self.srec = info(4)
        self.arena_bytes = info(5)
        self.state_words = info(6)
        self.max_actions = info(7)
        self.ready_cap = info(8)
        self.kernels_per_step = info(9)
        self.eval_rows = info(10)            # evaluation rows per game (a Player with PuctConfig.batch_size > 1)
        rows = n_games * self.eval_rows     # row g * eval_rows + k = request k of game g
        self.plane_shape = plane_shape
        self.plane_c_pad = plane_c_pad
        if plane_c_pad:
            # channels-last, channel count padded: logical [G, c_pad, H, W] over [G, H, W, c_pad] memory
            self._L.gz_puct_set_plane_layout.argtypes = [ctypes.c_void_p, ctypes.c_int]
            _lib.check(self._L.gz_puct_set_plane_layout(self._h, plane_c_pad))
            mem = torch.zeros((rows, plane_shape[1], plane_shape[2], plane_c_pad), dtype=plane_dtype, device=self.device)
            self.planes = mem.permute(0, 3, 1, 2)
        else:
            self.planes = torch.zeros((rows,) + tuple(plane_shape), dtype=plane_dtype, device=self.device)
        self.policy = torch.zeros((rows, self.policy_stride), dtype=torch.float32, device=self.device)
        self.value = torch.zeros((rows, self.n_rewards), dtype=torch.float32, device=self.device)
        self._fresh = False
        self.role_sizes = list(role_sizes)

    def __del__(self):
        h = getattr(self, '_h', None)
        if h and getattr(self, '_owns', False):
            self._L.gz_puct_destroy(h)
        self._h = None

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- the poll-equivalent on device tensors
    def join(self):
        """Waits (on the current stream) for the side-stream work of the last ``advance(..., defer_join=True)``."""
        _lib.check(self._L.gz_puct_join(self._h, self._stream()))

    def advance(self, policy=None, value=None, sub_iters=1, defer_join=False):
        """policy f32 [G, sum(role sizes)] (role rows concatenated), value f32 [G, n_rewards]: the
        network outputs for the planes returned by the previous call (None on the first call).
        With ``defer_join`` the step's side-stream work (re-rooting, goal values of terminal states) keeps
        running while the caller enqueues the network forward; ``join()`` must follow on the same stream.
        Returns planes [G, C, H, W] (bf16) — one row per game; rows of games not waiting for the
        network are to be ignored (their outputs are)."""
        fresh = policy is not None
        if fresh:
            assert policy.dtype == torch.float32 and policy.is_contiguous() and policy.shape == self.policy.shape
            assert value.dtype == torch.float32 and value.is_contiguous() and value.shape == self.value.shape
        code = {torch.float32: 0, torch.bfloat16: 1}[self.planes.dtype]
        fn = self._L.gz_puct_advance_async if defer_join else self._L.gz_puct_advance
        _lib.check(fn(self._h, policy.data_ptr() if fresh else None, value.data_ptr() if fresh else None,
                      1 if fresh else 0, sub_iters, self.planes.data_ptr(), code, self._stream()))
        return self.planes

    # ---- compacted network batches: only the rows that wait for the network are handed out
    def compact_requests(self, cap=None, count_pinned=None):
        """After ``advance``: gathers the planes of the (at most ``cap``) waiting rows, in ascending game order, into a
        dense batch and returns it (``[cap, C, H, W]`` view; rows beyond the number handed out are stale). Requests
        beyond ``cap`` stay pending. ``count_pinned``: a pinned int32 tensor that receives the number of waiting rows
        asynchronously (read it later to size the next batches)."""
        cap = self.n_games if cap is None else int(cap)
        if getattr(self, '_planes_c', None) is None:
            # same memory layout as the planes (channels-last padded: logical [G, c_pad, H, W] over [G, H, W, c_pad])
            if self.plane_c_pad:
                mem = torch.zeros((self.planes.shape[0], self.plane_shape[1], self.plane_shape[2], self.plane_c_pad),
                                  dtype=self.planes.dtype, device=self.device)
                self._planes_c = mem.permute(0, 3, 1, 2)
            else:
                self._planes_c = torch.zeros_like(self.planes, memory_format=torch.contiguous_format)
        row_bytes = self.planes[0].numel() * self.planes.element_size()
        _lib.check(self._L.gz_puct_compact_requests(
            self._h, cap, self.planes.data_ptr(), row_bytes, self._planes_c.data_ptr(),
            count_pinned.data_ptr() if count_pinned is not None else None, self._stream()))
        return self._planes_c[:cap]

    def advance_compact(self, policy_c, value_c, sub_iters=1, defer_join=False):
        """The step after ``compact_requests``: ``policy_c`` / ``value_c`` are the network's outputs for the rows of the
        compacted batch, in its order (at least as many rows as were handed out)."""
        assert policy_c.dtype == torch.float32 and policy_c.is_contiguous() and policy_c.shape[1] == self.policy_stride
        assert value_c.dtype == torch.float32 and value_c.is_contiguous() and value_c.shape[1] == self.n_rewards
        code = {torch.float32: 0, torch.bfloat16: 1}[self.planes.dtype]
        _lib.check(self._L.gz_puct_advance_compact(self._h, policy_c.data_ptr(), value_c.data_ptr(), sub_iters,
                                                   self.planes.data_ptr(), code, self._stream(), 1 if defer_join else 0))
        return self.planes

    def advance_timed(self, policy=None, value=None):
        """One engine step with CUDA events between its launches (measurement aid); returns
        (planes, {kernel: ms})."""
        fresh = policy is not None
        code = {torch.float32: 0, torch.bfloat16: 1}[self.planes.dtype]
        ms = (ctypes.c_float * 5)()
        self._L.gz_puct_advance_timed.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                  ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                                  ctypes.POINTER(ctypes.c_float)]
        _lib.check(self._L.gz_puct_advance_timed(self._h, policy.data_ptr() if fresh else None,
                                                 value.data_ptr() if fresh else None, 1 if fresh else 0,
                                                 self.planes.data_ptr(), code, self._stream(), ms))
        names = ['puct_pre_kernel', 'puct_reroot_kernel', 'pn_kernel', 'puct_post_kernel', 'encode_kernel']
        return self.planes, dict(zip(names, [float(x) for x in ms]))

    def synth_nn(self, planes=None, seed=1):
        """Deterministic synthetic network (include/gz_synth.h) into self.policy / self.value."""
        planes = self.planes if planes is None else planes
        rs = (ctypes.c_int * len(self.role_sizes))(*self.role_sizes)
        code = {torch.float32: 0, torch.bfloat16: 1}[planes.dtype]
        _lib.check(self._L.gz_synth_nn(planes.data_ptr(), code, planes.shape[0], self.total_size, seed,
                                       len(self.role_sizes), rs, self.n_rewards, self.policy.data_ptr(),
                                       self.value.data_ptr(), self._stream()))
        return self.policy, self.value

    def run_synthetic(self, steps, seed=1, sub_iters=1):
        for _ in range(steps):
            if self._fresh:
                self.advance(self.policy, self.value, sub_iters)
            else:
                self.advance(None, None, sub_iters)
            self.synth_nn(seed=seed)
            self._fresh = True

    # ---- results
    def totals(self):
        out = (ctypes.c_double * len(TOTAL_KEYS))()
        _lib.check(self._L.gz_puct_totals(self._h, out, self._stream()))
        return dict(zip(TOTAL_KEYS, [float(x) for x in out]))

    def fetch_sample_records(self, cap=None):
        cap = cap or self.ready_cap
        buf = np.zeros((cap, self.srec), dtype=np.int32)
        n = self._L.gz_puct_fetch_samples(self._h, buf.ctypes.data, cap, self._stream())
        if n < 0:
            _lib.check(int(n))
        return buf[:n]

    def records_to_samples(self, recs, identifier='b200'):
        """int32 records -> list of dicts with the keys of ``Supervisor.fetch_samples``
        (``src/cpp/pyobjects/supervisor_impl.cpp:75-118`` == ``datadesc.Sample`` fields)."""
        from . import tables
        nb = self.n_bases
        sw = self.state_words
        R = len(self.role_sizes)
        out = []
        for r in recs:
            f = r.view(np.float32)
            nc = int(r[SR_N_CHILDREN])
            lead = int(r[SR_LEAD])
            state = tables.unpack_state_bits(r[SR_HEADER:SR_HEADER + sw].view(np.uint32), nb)
            prevs = []
            if r[SR_N_PREV]:
                prevs.append([int(x) for x in tables.unpack_state_bits(
                    r[SR_HEADER + sw:SR_HEADER + 2 * sw].view(np.uint32), nb)])
            pol = r[SR_HEADER + 2 * sw:SR_HEADER + 2 * sw + 2 * nc].reshape(nc, 2)
            policies = []
            for ri in range(R):
                if ri == lead:
                    policies.append([(int(m), float(p)) for m, p in zip(pol[:, 0], pol[:, 1].view(np.float32))])
                else:
                    policies.append([(int(r[SR_OTHER_MOVES + ri]), 1.0)])
            out.append(dict(
                state=[int(x) for x in state], prev_states=prevs, policies=policies,
                final_score=[float(f[SR_FINAL_SCORE + i]) for i in range(R)],
                depth=int(r[SR_DEPTH]), game_length=int(r[SR_GAME_LENGTH]),
                match_identifier='%s_%d_%d' % (identifier, int(r[SR_SLOT]), int(r[SR_MATCH])),
                has_resigned=bool(r[SR_HAS_RESIGNED]), resign_false_positive=bool(r[SR_RESIGN_FP]),
                starting_sample_depth=int(r[SR_START_DEPTH]),
                resultant_puct_score=[float(f[SR_PUCT_SCORE + i]) for i in range(R)],
                resultant_puct_visits=int(r[SR_VISITS])))
        return out

    def set_start_states(self, states):
        """states int32 CUDA tensor [n_games, state_words]: the next game of every slot starts there (depth 0)."""
        assert states.is_cuda and states.dtype == torch.int32 and states.is_contiguous()
        assert tuple(states.shape) == (self.n_games, self.state_words)
        _lib.check(self._L.gz_puct_set_start_states(self._h, states.data_ptr(), self._stream()))

    def set_step_limits(self, max_iters=16, max_work=1 << 20):
        self._L.gz_puct_set_step_limits.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        _lib.check(self._L.gz_puct_set_step_limits(self._h, int(max_iters), int(max_work)))

    def clear_unique_states(self):
        _lib.check(self._L.gz_puct_clear_unique_states(self._h, self._stream()))

    # ---- player mode
    def command(self, game, cmd, arg=0, evals=0, state=None):
        sp = None
        if state is not None:
            state = np.ascontiguousarray(state, dtype=np.uint32)
            sp = state.ctypes.data
        _lib.check(self._L.gz_puct_command(self._h, game, cmd, arg, evals, sp, self._stream()))
        if state is not None:
            torch.cuda.current_stream(self.device).synchronize()   # host buffer must outlive the copy

    def snapshot(self, game=0, max_children=None):
        mc = max_children or self.max_actions
        oi = np.zeros(16, dtype=np.int32)
        of = np.zeros(8, dtype=np.float32)
        moves = np.zeros(mc, dtype=np.int32)
        trav = np.zeros(mc, dtype=np.int32)
        prob = np.zeros(mc, dtype=np.float32)
        score = np.zeros(mc, dtype=np.float32)
        nxt = np.zeros(mc, dtype=np.float32)
        n = self._L.gz_puct_snapshot(self._h, game, mc, oi.ctypes.data, of.ctypes.data, moves.ctypes.data,
                                     trav.ctypes.data, prob.ctypes.data, score.ctypes.data, nxt.ctypes.data,
                                     self._stream())
        if n < 0:
            _lib.check(n)
        return dict(cmd_done=int(oi[0]), phase=int(oi[1]), error=int(oi[2]), choice=int(oi[3]),
                    nodes=int(oi[4]), children=int(oi[5]), root_visits=int(oi[6]), lead_role=int(oi[7]),
                    other_moves=[int(x) for x in oi[9:9 + len(self.role_sizes)]],
                    root_score=float(of[0]), choice_scores=[float(x) for x in of[1:1 + len(self.role_sizes)]],
                    moves=moves[:n].copy(), traversals=trav[:n].copy(),
                    policy_prob=prob[:n].copy(), scores=score[:n].copy(), next_prob=nxt[:n].copy())
