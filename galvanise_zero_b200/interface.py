"""``ggpzero_interface`` — the reference's extension-module surface, re-hosted on libgz_b200.

The reference module (``src/cpp/ggpzero_interface.cpp:50-96``) exports three factories —
``GdlBasesTransformer``, ``Player``, ``Supervisor`` — plus ``buf_to_tuple_reverse_bytes`` and the
exception ``GGPZeroInterfaceError``. The same names, argument orders and method names are kept here;
each object forwards to the C-ABI of ``include/gz_b200.h``. New, additive methods (``encode`` on
device tensors, ``poll_device``) carry the north-star requirement that planes never round-trip
through the host.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import lib as _lib

GGPZeroInterfaceError = _lib.GzError


def buf_to_tuple_reverse_bytes(buf) -> tuple:
    """bytes -> tuple of bools, MSB first per byte (inverse of ``np.packbits``);
    reference: ``src/cpp/ggpzero_interface.cpp:22-46``."""
    a = np.frombuffer(bytes(buf), dtype=np.uint8)
    return tuple(bool(x) for x in np.unpackbits(a))


def _declare(L):
    if getattr(L, '_gz_tr_declared', False):
        return
    vp, ci, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    L.gz_transformer_create.argtypes = [ci, ci, ci, ci, ci, ctypes.POINTER(ci), ci, ctypes.POINTER(vp)]
    L.gz_transformer_destroy.argtypes = [vp]
    L.gz_transformer_add_board_base.argtypes = [vp, ci, ci]
    L.gz_transformer_add_control_base.argtypes = [vp, ci, ci, ctypes.c_float]
    L.gz_transformer_finalize.argtypes = [vp, ci, ci]
    L.gz_transformer_total_size.argtypes = [vp]
    L.gz_encode.argtypes = [vp, i64, ci, vp, vp, vp, ci, vp, ci, vp]
    L.gz_encode_layout.argtypes = [vp, i64, ci, vp, vp, vp, ci, vp, ci, ci, vp]
    L._gz_tr_declared = True


class GdlBasesTransformer(object):
    """C-side transformer (reference: ``src/cpp/pyobjects/gdltransformer_impl.cpp:186-219``)."""

    def __init__(self, channel_size, channels_per_state, num_control_channels, num_prev_states,
                 num_rewards, policy_sizes):
        self._L = _lib.lib()
        _declare(self._L)
        self.channel_size = int(channel_size)
        self.channels_per_state = int(channels_per_state)
        self.num_control_channels = int(num_control_channels)
        self.num_prev_states = int(num_prev_states)
        self.num_rewards = int(num_rewards)
        self.policy_sizes = [int(x) for x in policy_sizes]
        arr = (ctypes.c_int * len(self.policy_sizes))(*self.policy_sizes)
        h = ctypes.c_void_p()
        _lib.check(self._L.gz_transformer_create(self.channel_size, self.channels_per_state,
                                                 self.num_control_channels, self.num_prev_states,
                                                 self.num_rewards, arr, len(self.policy_sizes),
                                                 ctypes.byref(h)))
        self._h = h
        self._finalized_for = None
        self.board_bases = []       # kept for introspection / oracle construction in tests
        self.control_bases = []

    def __del__(self):
        h = getattr(self, '_h', None)
        if h:
            self._L.gz_transformer_destroy(h)
            self._h = None

    # ---- builder methods (same names as the reference)
    def add_board_base(self, base_idx, buf_offset):
        _lib.check(self._L.gz_transformer_add_board_base(self._h, int(base_idx), int(buf_offset)))
        self.board_bases.append((int(base_idx), int(buf_offset)))

    def add_control_base(self, base_idx, channel_id, value):
        _lib.check(self._L.gz_transformer_add_control_base(self._h, int(base_idx), int(channel_id),
                                                           float(value)))
        self.control_bases.append((int(base_idx), int(channel_id), float(value)))

    def total_size(self):
        return int(self._L.gz_transformer_total_size(self._h))

    @property
    def num_channels(self):
        return self.channels_per_state * (self.num_prev_states + 1) + self.num_control_channels

    # ---- device side
    def finalize(self, n_bases, device=0):
        if self._finalized_for is None:
            _lib.check(self._L.gz_transformer_finalize(self._h, int(n_bases), int(device)))
            self._finalized_for = (int(n_bases), int(device))
        return self

    def encode(self, states, prev_states=None, prev_count=None, out=None, dtype=None, c_pad=0):
        """states int32 [n, sw] (CUDA) -> planes [n, total_size] (bf16 default, or float32).
        prev_states int32 [n, max_prev, sw], nearest ancestor first; prev_count uint8 [n]."""
        import torch
        assert self._finalized_for is not None, 'call finalize(n_bases, device) first'
        assert states.is_cuda and states.dtype == torch.int32 and states.is_contiguous()
        n, sw = states.shape
        if dtype is None:
            dtype = out.dtype if out is not None else torch.bfloat16
        row = self.channel_size * c_pad if c_pad else self.total_size()   # c_pad: [n][H*W][c_pad] rows
        if out is None:
            out = torch.empty((n, row), dtype=dtype, device=states.device)
        assert out.is_contiguous() and out.shape[0] == n and out[0].numel() == row
        code = {torch.float32: 0, torch.bfloat16: 1}[out.dtype]
        max_prev = 0
        pp = pc = None
        if prev_states is not None:
            assert prev_states.is_cuda and prev_states.dtype == torch.int32 and prev_states.is_contiguous()
            assert prev_states.dim() == 3 and prev_states.shape[0] == n and prev_states.shape[2] == sw
            max_prev = prev_states.shape[1]
            pp = prev_states.data_ptr()
            if prev_count is not None:
                assert prev_count.dtype == torch.uint8 and prev_count.shape == (n,)
                pc = prev_count.data_ptr()
        stream = ctypes.c_void_p(torch.cuda.current_stream(states.device).cuda_stream)
        _lib.check(self._L.gz_encode_layout(self._h, n, sw, states.data_ptr(), pp, pc, max_prev,
                                            out.data_ptr(), code, c_pad, stream))
        return out


# ------------------------------------------------------------------------------------------------
class _Manager(object):
    """One engine + its stream — the device counterpart of a ``SelfPlayManager``
    (``src/cpp/selfplaymanager.h:20-114``). The hand-off itself (pinned buffers, copy-in, engine step,
    compaction of the waiting rows, copy-out) is native: ``gz_puct_host_poll_submit`` / ``_wait``."""

    def __init__(self, engine, stream):
        e = self.engine = engine
        self.stream = stream                     # None: the caller's current stream (inline mode)
        self.n_req = 0                           # rows handed out and not yet answered
        L = e._L
        L.gz_puct_host_poll_submit.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p),
                                               ctypes.c_int, ctypes.c_void_p]
        L.gz_puct_host_poll_wait.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.POINTER(ctypes.c_float))]
        self._n_arrays = len(e.role_sizes) + 1
        self._ptrs = (ctypes.c_void_p * self._n_arrays)()
        self._planes_ptr = ctypes.POINTER(ctypes.c_float)()
        self._planes_np = None

    def _ctx(self):
        import contextlib
        import torch
        return torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()

    def _stream_ptr(self):
        import torch
        s = self.stream if self.stream is not None else torch.cuda.current_stream(self.engine.device)
        return ctypes.c_void_p(s.cuda_stream)

    def launch(self, n_predictions, arrays=None):
        """Enqueues, without waiting: predictions -> engine rows, one engine step, the compaction of
        the rows that now wait for the network and the copy of their planes to pinned host memory."""
        e = self.engine
        if n_predictions:
            for i in range(self._n_arrays):
                self._ptrs[i] = arrays[i].ctypes.data
        _lib.check(e._L.gz_puct_host_poll_submit(e._h, int(n_predictions), self._ptrs if n_predictions else None,
                                                 self._n_arrays if n_predictions else 0, self._stream_ptr()))

    def wait(self):
        e = self.engine
        n = e._L.gz_puct_host_poll_wait(e._h, ctypes.byref(self._planes_ptr))
        if n < 0:
            _lib.check(n)
        if self._planes_np is None:
            self._planes_np = np.ctypeslib.as_array(self._planes_ptr, shape=(e.n_games * e.total_size,))
        self.n_req = n
        return n

    def planes(self):
        return self._planes_np[:self.n_req * self.engine.total_size]


class Supervisor(object):
    """``ggpzero_interface.Supervisor(sm, transformer, batch_size, identifier)``
    (``src/cpp/pyobjects/supervisor_impl.cpp:215-237``) on the device engine.

    ``poll`` keeps the reference's numpy contract (``src/cpp/pyobjects/common.cpp:160-218``): it takes
    the predictions for the rows handed out by the previous poll and returns a 1-D float32 array of
    ``n_req * total_size`` planes (or None when there is nothing to evaluate). ``poll_device`` is the
    additive zero-copy variant on CUDA tensors.

    ``start_self_play(num_workers, conf)``: ``num_workers <= 0`` is the reference's inline mode — one
    manager holding ``batch_size`` games, strictly serial. ``num_workers > 0`` is its worker mode
    (``src/cpp/supervisor.cpp:79-99``: two managers per worker so that the search of one overlaps the
    network call of the other): here, likewise, two managers of ``batch_size`` games each on their own
    CUDA streams — while the caller runs the network on the batch of one manager, the other manager's engine
    step runs on the device. Batches alternate between the managers, as ready managers do in the
    reference (``supervisor.cpp:129-167``)."""

    def __init__(self, sm, c_transformer, batch_size, identifier, device=0, seed=0x5EED0000,
                 node_cap=4096, child_cap=None):
        self.sm = sm
        self.game = sm if isinstance(sm, str) else getattr(sm, 'game', None)
        if self.game is None:
            raise GGPZeroInterfaceError('Supervisor needs a state machine created by name')
        self.c_transformer = c_transformer
        self.batch_size = int(batch_size)
        self.identifier = identifier
        self.device = device
        self.seed = seed
        self.node_cap, self.child_cap = node_cap, child_cap
        self.engine = None
        self.managers = []
        self._cur = None            # manager whose planes were handed out by the last poll
        self._primed = False
        # additive (measurement aid): fn(engine, manager_index) -> int32 CUDA tensor [games, state_words]; the
        # first game of every slot then starts there instead of the initial state (PuctEngine.set_start_states)
        self.start_states_fn = None

    def set_num_workers(self, n):
        """Not in the reference's method table although its Python wrapper calls it
        (``cppinterface.py:197-198``); accepted and ignored — the device engine has no worker threads."""

    def start_self_play(self, num_workers, conf_dict):
        from .engine import PuctEngine
        import torch
        n_man = 2 if num_workers > 0 else 1
        sizes = [self.batch_size] * n_man
        self.managers = []
        dsm = self.sm if not isinstance(self.sm, str) else None
        for k, g in enumerate(sizes):
            eng = PuctEngine(self.game, conf_dict, n_games=g, mode=0, device=self.device,
                             seed=self.seed + 104729 * k, node_cap=self.node_cap, child_cap=self.child_cap,
                             c_transformer=self.c_transformer, plane_dtype=torch.float32, sm=dsm)
            dsm = eng.sm                                      # managers share the device program
            if self.start_states_fn is not None:
                eng.set_start_states(self.start_states_fn(eng, k))
            stream = torch.cuda.Stream(device=eng.device) if n_man > 1 else None
            self.managers.append(_Manager(eng, stream))
        self.engine = self.managers[0].engine
        torch.cuda.synchronize(self.engine.device)
        self._cur = None
        self._primed = False

    # -- zero-copy path (inline mode)
    def poll_device(self, policy=None, value=None, sub_iters=1):
        return self.engine.advance(policy, value, sub_iters)

    # -- reference-shaped numpy path
    def poll(self, n_predictions, arrays):
        """Returns a 1-D float32 numpy array VIEWING engine-owned pinned host memory (valid until the
        next poll that returns the same manager's batch, as in the reference: common.cpp:204-209), or
        None when nothing needs evaluating."""
        if not self.managers:
            raise GGPZeroInterfaceError('start_self_play() has not been called')
        if not (0 <= n_predictions <= self.batch_size):
            raise GGPZeroInterfaceError('n_predictions out of range')
        M = self.managers
        if not self._primed:
            if n_predictions != 0:
                raise GGPZeroInterfaceError('first poll must pass n_predictions == 0')
            for m in M:
                m.launch(0)
            self._primed = True
            nxt = 0
        else:
            cur = self._cur
            m = M[cur]
            if n_predictions != m.n_req:
                raise GGPZeroInterfaceError('expected %d predictions, got %d' % (m.n_req, n_predictions))
            if n_predictions:
                e = m.engine
                R = len(e.role_sizes)
                if len(arrays) != R + 1:
                    raise GGPZeroInterfaceError('expected %d policy arrays and one value array' % R)
                widths = list(e.role_sizes) + [e.n_rewards]
                for a, w in zip(arrays, widths):
                    if not isinstance(a, np.ndarray) or a.dtype != np.float32 or not a.flags['C_CONTIGUOUS']:
                        raise GGPZeroInterfaceError('predictions must be C-contiguous float32 numpy arrays')
                    if a.size != n_predictions * w:
                        raise GGPZeroInterfaceError('prediction array holds %d values, expected %d x %d'
                                                    % (a.size, n_predictions, w))
            m.n_req = 0
            m.launch(n_predictions, arrays)
            nxt = (cur + 1) % len(M)
        # hand out the next ready batch; a manager with nothing to evaluate is stepped again
        for _ in range(64 * len(M)):
            m = M[nxt]
            if m.wait() > 0:
                self._cur = nxt
                return m.planes()
            m.launch(0)
            nxt = (nxt + 1) % len(M)
        self._primed = False
        self._cur = None
        return None

    def totals(self):
        out = None
        for m in self.managers:
            with m._ctx():                         # ordered behind the manager's queued engine step
                t = m.engine.totals()
            out = t if out is None else {k: (max(out[k], t[k]) if k.startswith('max_') else out[k] + t[k])
                                         for k in out}
        return out

    def check_health(self):
        """Raises when game slots were lost (halted) or finished samples were dropped (fetch_samples called too
        rarely for the ready buffer); the reference's sample vector is unbounded and its games never stop."""
        t = self.totals()
        if t['halted'] > 0 or t['dropped_samples'] > 0:
            raise GGPZeroInterfaceError('device self-play degraded: %d game slots halted (first error %d), %d samples '
                                        'dropped' % (t['halted'], t['first_error'], t['dropped_samples']))
        return t

    def fetch_samples(self):
        res = []
        for k, m in enumerate(self.managers):
            with m._ctx():
                recs = m.engine.fetch_sample_records()
            if len(recs):
                ident = self.identifier if len(self.managers) == 1 else '%s_m%d' % (self.identifier, k)
                res.extend(m.engine.records_to_samples(recs, ident))
        return res or None

    def add_unique_state(self, bs):
        """The reference's Python wrapper no-ops this call as well (``cppinterface.py:218-219``)."""

    def clear_unique_states(self):
        for m in self.managers:
            with m._ctx():
                m.engine.clear_unique_states()


# ------------------------------------------------------------------------------------------------
class Player(object):
    """``ggpzero_interface.Player(sm, transformer, puct_conf_dict)``
    (``src/cpp/pyobjects/player_impl.cpp:8-129,183-201``, ``src/cpp/player.cpp``) on the device engine
    in player mode (one tree).

    As in the reference, ``player_apply_move`` / ``player_move`` only *enqueue* work
    (``player.cpp:59-103``); it is carried out by the following ``poll`` calls, which hand out the
    planes of the leaf to evaluate and return ``None`` once the queue has drained.

    Differences from the reference: one outstanding evaluation per tree whatever
    ``batch_size`` says; ``end_time`` / ``think_time`` wall-clock limits are not consulted (the search
    stops on its evaluation budget and convergence rules); ``player_balance_moves`` is accepted and
    does nothing; ``player_tree_debug`` reports no variation lines."""

    CMD_RESET, CMD_APPLY, CMD_SEARCH = 1, 2, 3

    def __init__(self, sm, c_transformer, puct_conf_dict, device=0, seed=0x5EED0001, node_cap=1 << 16,
                 child_cap=None):
        from .engine import PuctEngine
        from .defs import confs
        import attr
        import torch
        self.sm = sm
        self.game = sm if isinstance(sm, str) else getattr(sm, 'game', None)
        if self.game is None:
            raise GGPZeroInterfaceError('Player needs a state machine created by name')
        if not isinstance(puct_conf_dict, dict):
            raise GGPZeroInterfaceError('puct config must be a dict (attr.asdict(PUCTEvaluatorConfig))')
        self.conf = dict(attr.asdict(confs.PUCTEvaluatorConfig()))
        self.conf.update(puct_conf_dict)
        sp = attr.asdict(confs.selfplay_config_template())
        sp['puct_config'] = dict(self.conf)
        sp['run_to_end_puct_config'] = dict(self.conf)
        self._sp = sp
        self.c_transformer = c_transformer
        self.engine = PuctEngine(self.game, sp, n_games=1, mode=PuctEngine.MODE_PLAYER, device=device, seed=seed,
                                 node_cap=node_cap, child_cap=child_cap, c_transformer=c_transformer,
                                 plane_dtype=torch.float32,
                                 sm=self.sm if not isinstance(self.sm, str) else None)
        e = self.engine
        self._host_planes = torch.empty((1, e.total_size), dtype=torch.float32).pin_memory()
        self._queue = []            # pending (kind, payload) commands, reference: scheduler runnables
        self._busy = False          # a command is in the engine's mailbox
        self._current = None
        self._waiting = 0           # rows handed out by the last poll
        self._first_play = False
        self._game_depth = 0
        self._choice = None         # (snapshot) of the last finished search
        self._fresh = False

    # ---- reference method table
    def player_reset(self, game_depth):
        self._game_depth = int(game_depth)
        self._first_play = True
        self._queue = []
        self._choice = None

    def player_update_config(self, think_time, converged_visits, verbose):
        from .engine import _puct_array, _sp_array
        self.conf['think_time'] = float(think_time)
        self.conf['converged_visits'] = int(converged_visits)
        self.conf['verbose'] = bool(verbose)
        self._sp['puct_config'] = dict(self.conf)
        self._sp['run_to_end_puct_config'] = dict(self.conf)
        e = self.engine
        _lib.check(e._L.gz_puct_set_config(e._h, _sp_array(self._sp), _puct_array(self.conf),
                                           _puct_array(self.conf)))

    def _joint_move_list(self, jm):
        if hasattr(jm, 'get') and hasattr(jm, 'get_count'):
            return [int(jm.get(i)) for i in range(jm.get_count())]
        return [int(x) for x in jm]

    def _state_words(self, bs):
        from . import tables
        if hasattr(bs, 'to_list'):
            bits = np.asarray(bs.to_list(), dtype=np.uint8)
            return tables.pack_state_bits(bits, self.engine.state_words)
        return np.ascontiguousarray(bs, dtype=np.uint32)

    def player_apply_move(self, joint_move):
        moves = self._joint_move_list(joint_move)
        if self._first_play:
            self._first_play = False
            self._queue.append(('reset', None))           # establishRoot(nullptr): the initial state
        self._queue.append(('apply', moves))

    def player_move(self, basestate, evaluations, end_time):
        self._choice = None
        if self._first_play:
            self._first_play = False
            self._queue.append(('reset', self._state_words(basestate)))
        self._queue.append(('search', int(evaluations)))

    def player_get_move(self, lead_role_index):
        """-> (legal index of ``lead_role_index``'s move, score of the chosen child or -1, node count)."""
        s = self._choice
        if s is None or s['choice'] < 0 or len(s['moves']) == 0:
            return (-1, -1.0, -1)
        ch = s['choice']
        if lead_role_index == s['lead_role']:
            move = int(s['moves'][ch])
        else:
            move = int(s['other_moves'][lead_role_index])
        # the chosen child's score FOR THE ROLE ASKED ABOUT (player.cpp:105-119), not for the root's lead role
        return (move, float(s['choice_scores'][lead_role_index]), int(s['nodes']))

    def player_balance_moves(self, max_count):
        """Accepted for API parity; the device engine does not rebalance first moves
        (``src/cpp/puct/evaluator.cpp:282-339`` is an analysis aid of the tournament tooling)."""

    def player_tree_debug(self, max_count):
        """-> tuple of (score, lead_role_index, move_index, variation) for the root's most visited
        expanded children (``src/cpp/puct/node.cpp:400-440``); variation lines are not reported."""
        s = self.engine.snapshot(0)
        order = sorted(range(len(s['moves'])), key=lambda i: -int(s['traversals'][i]))
        out = []
        for i in order[:max(0, int(max_count))]:
            if s['scores'][i] < 0:
                continue
            out.append((float(s['scores'][i]), int(s['lead_role']), int(s['moves'][i]), ()))
        return tuple(out)

    # ---- the poll contract (src/cpp/player.cpp:155-173, src/cpp/pyobjects/common.cpp:160-218)
    def _issue_next(self):
        e = self.engine
        kind, payload = self._queue.pop(0)
        self._current = kind
        if kind == 'reset':
            e.command(0, self.CMD_RESET, arg=self._game_depth, state=payload)
        elif kind == 'apply':
            s = e.snapshot(0, 0)
            lead = s['lead_role'] if s['lead_role'] >= 0 else 0
            e.command(0, self.CMD_APPLY, arg=payload[lead])
            self._game_depth += 1
        else:
            e.command(0, self.CMD_SEARCH, evals=payload)
        self._busy = True

    def poll(self, n_predictions, arrays):
        import torch
        e = self.engine
        if n_predictions != self._waiting:
            raise GGPZeroInterfaceError('expected %d predictions, got %d' % (self._waiting, n_predictions))
        fresh = False
        if n_predictions:
            R = len(e.role_sizes)
            if len(arrays) != R + 1:
                raise GGPZeroInterfaceError('expected %d policy arrays and one value array' % R)
            off = 0
            row = np.zeros((1, e.policy_stride), dtype=np.float32)
            for r in range(R):
                a = arrays[r]
                if not isinstance(a, np.ndarray) or a.dtype != np.float32 or not a.flags['C_CONTIGUOUS']:
                    raise GGPZeroInterfaceError('predictions must be C-contiguous float32 numpy arrays')
                row[0, off:off + e.role_sizes[r]] = a.reshape(-1)[:e.role_sizes[r]]
                off += e.role_sizes[r]
            e.policy.copy_(torch.from_numpy(row))
            e.value.copy_(torch.from_numpy(np.ascontiguousarray(arrays[R], dtype=np.float32).reshape(1, e.n_rewards)))
            fresh = True
        self._waiting = 0
        guard = 0
        while True:
            if not self._busy:
                if not self._queue:
                    return None
                self._issue_next()
            e.advance(e.policy if fresh else None, e.value if fresh else None)
            fresh = False
            if not hasattr(self, '_need_buf'):
                self._need_buf = torch.zeros(1, dtype=torch.uint8, device=e.device)
                e._L.gz_puct_copy_need_eval.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
            _lib.check(e._L.gz_puct_copy_need_eval(e._h, self._need_buf.data_ptr(), e._stream()))
            need = int(self._need_buf[0].item())                       # device sync
            if need:
                self._host_planes.copy_(e.planes.reshape(1, -1).to(torch.float32), non_blocking=True)
                torch.cuda.current_stream(e.device).synchronize()
                self._waiting = 1
                return self._host_planes.numpy().reshape(-1)
            guard += 1
            if guard % 4 == 0 or guard < 3:
                s = e.snapshot(0, 0)
                if s['error'] or s['cmd_done'] < 0:
                    raise GGPZeroInterfaceError('device player failed: error %d, cmd_done %d' % (s['error'], s['cmd_done']))
                if s['cmd_done'] == 1:
                    self._busy = False
                    if self._current == 'search':
                        self._choice = e.snapshot(0)
            if guard > 10000000:
                raise GGPZeroInterfaceError('device player did not finish')
