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
def _declare_host(L):
    if getattr(L, '_gz_host_declared', False):
        return
    vp, ci, i64, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64
    dp, fp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float)
    L.gz_supervisor_create.argtypes = [vp, vp, ci, ctypes.c_char_p, ci, u64, ci, ci, ctypes.POINTER(vp)]
    L.gz_supervisor_destroy.argtypes = [vp]
    L.gz_supervisor_start_self_play.argtypes = [vp, ci, dp, dp, dp]
    L.gz_supervisor_info.argtypes = [vp, ci]
    L.gz_supervisor_info.restype = i64
    L.gz_supervisor_engine.argtypes = [vp, ci]
    L.gz_supervisor_engine.restype = vp
    L.gz_supervisor_stream.argtypes = [vp, ci]
    L.gz_supervisor_stream.restype = vp
    L.gz_supervisor_poll.argtypes = [vp, ci, ctypes.POINTER(vp), ci, ctypes.POINTER(fp)]
    L.gz_supervisor_fetch_samples.argtypes = [vp, vp, i64, vp]
    L.gz_supervisor_fetch_samples.restype = i64
    L.gz_supervisor_clear_unique_states.argtypes = [vp]
    L.gz_supervisor_totals.argtypes = [vp, dp]
    L.gz_player_create.argtypes = [vp, vp, dp, ci, u64, ci, ci, ctypes.POINTER(vp)]
    L.gz_player_destroy.argtypes = [vp]
    L.gz_player_engine.argtypes = [vp]
    L.gz_player_engine.restype = vp
    L.gz_player_info.argtypes = [vp, ci]
    L.gz_player_info.restype = i64
    L.gz_player_reset.argtypes = [vp, ci]
    L.gz_player_update_config.argtypes = [vp, ctypes.c_double, ci, ci]
    L.gz_player_apply_move.argtypes = [vp, ctypes.POINTER(ci), ci]
    L.gz_player_move.argtypes = [vp, vp, ci, ctypes.c_double]
    L.gz_player_get_move.argtypes = [vp, ci, ctypes.POINTER(ci), fp, ctypes.POINTER(ci)]
    L.gz_player_balance_moves.argtypes = [vp, ci]
    L.gz_player_tree_debug.argtypes = [vp, ci, vp, vp, vp]
    L.gz_player_tree_debug_lines.argtypes = [vp, ci, vp, vp, vp, vp, vp]
    L.gz_player_poll.argtypes = [vp, ci, ctypes.POINTER(vp), ci, ctypes.POINTER(fp)]
    L.gz_player_last_search.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ci]
    L._gz_host_declared = True


def _resolve_sm(sm, device):
    """sm: a game name, a DeviceStateMachine, a ggplib-shaped StateMachine (ggplib_api) or its ``sm_ptr`` — the
    reference passes the state machine as a pointer cast to an int (``cppinterface.py:12-30``); here the pointer is
    the ``gz_propnet*`` of the device program. Returns (DeviceStateMachine or None, handle)."""
    from .statemachine import DeviceStateMachine
    if isinstance(sm, str):
        dsm = DeviceStateMachine(sm, device=device, goals_mode=_lib.GOALS_TERMINAL_ONLY)
        return dsm, dsm._h
    if isinstance(sm, int):
        return None, ctypes.c_void_p(sm)
    dev = getattr(sm, 'dev', sm)                 # ggplib_api.StateMachine keeps its DeviceStateMachine in .dev
    if getattr(dev, '_h', None) is None:
        raise GGPZeroInterfaceError('not a state machine: %r' % (sm,))
    return dev, dev._h


def _check_predictions(arrays, n, widths):
    if len(arrays) != len(widths):
        raise GGPZeroInterfaceError('expected %d policy arrays and one value array' % (len(widths) - 1))
    for a, w in zip(arrays, widths):
        if not isinstance(a, np.ndarray) or a.dtype != np.float32 or not a.flags['C_CONTIGUOUS']:
            raise GGPZeroInterfaceError('predictions must be C-contiguous float32 numpy arrays')
        if a.size != n * w:
            raise GGPZeroInterfaceError('prediction array holds %d values, expected %d x %d' % (a.size, n, w))


class _Manager(object):
    """View of one of the native supervisor's managers (one engine + its stream) — the device counterpart of a
    ``SelfPlayManager`` (``src/cpp/selfplaymanager.h:20-114``)."""

    def __init__(self, engine, stream_ptr):
        import torch
        self.engine = engine
        self.stream = torch.cuda.ExternalStream(stream_ptr, device=engine.device)

    def _ctx(self):
        import torch
        return torch.cuda.stream(self.stream)


class Supervisor(object):
    """``ggpzero_interface.Supervisor(sm, transformer, batch_size, identifier)``
    (``src/cpp/pyobjects/supervisor_impl.cpp:215-237``): a ctypes shell over the native ``gz_supervisor``
    (``csrc/gz_host.cu``), which holds the managers, their streams and the hand-over logic.

    ``poll`` keeps the reference's numpy contract (``src/cpp/pyobjects/common.cpp:160-218``): it takes
    the predictions for the rows handed out by the previous poll and returns a 1-D float32 array of
    ``n_req * total_size`` planes (or None when there is nothing to evaluate). ``poll_device`` is the
    additive zero-copy variant on CUDA tensors.

    ``start_self_play(num_workers, conf)``: ``num_workers <= 0`` is the reference's inline mode — one
    manager holding ``batch_size`` games, strictly serial. ``num_workers > 0`` is its worker mode
    (``src/cpp/supervisor.cpp:79-99``: two managers per worker so that the search of one overlaps the
    network call of the other): two managers of ``batch_size`` games each on their own CUDA streams."""

    def __init__(self, sm, c_transformer, batch_size, identifier, device=0, seed=0x5EED0000,
                 node_cap=4096, child_cap=None):
        self._L = _lib.lib()
        _declare_host(self._L)
        self.sm, self._pn = _resolve_sm(sm, device)
        self.c_transformer = c_transformer
        self.batch_size = int(batch_size)
        self.identifier = identifier
        self.device = device
        L = self._L
        self.n_bases = int(L.gz_propnet_info(self._pn, _lib.INFO_N_BASES))
        n_roles = int(L.gz_propnet_info(self._pn, _lib.INFO_N_ROLES))
        self.role_sizes = [int(L.gz_propnet_info(self._pn, _lib.INFO_ROLE_SIZE0 + r)) for r in range(n_roles)]
        c_transformer.finalize(self.n_bases, device)
        h = ctypes.c_void_p()
        _lib.check(L.gz_supervisor_create(self._pn, c_transformer._h, self.batch_size, str(identifier).encode(), device,
                                          seed, node_cap, child_cap or 0, ctypes.byref(h)))
        self._h = h
        self.engine = None
        self.managers = []
        self._ptrs = (ctypes.c_void_p * (n_roles + 1))()
        self._planes_ptr = ctypes.POINTER(ctypes.c_float)()
        self._views = {}
        # additive (measurement aid): fn(engine, manager_index) -> int32 CUDA tensor [games, state_words]; the
        # first game of every slot then starts there instead of the initial state (PuctEngine.set_start_states)
        self.start_states_fn = None

    def __del__(self):
        h = getattr(self, '_h', None)
        if h:
            self.managers = []
            self.engine = None
            self._L.gz_supervisor_destroy(h)
            self._h = None

    def set_num_workers(self, n):
        """Not in the reference's method table although its Python wrapper calls it
        (``cppinterface.py:197-198``); accepted and ignored — the device engine has no worker threads."""

    def start_self_play(self, num_workers, conf_dict):
        from .engine import PuctEngine, _puct_array, _sp_array
        import torch
        L = self._L
        self.managers = []
        _lib.check(L.gz_supervisor_start_self_play(self._h, int(num_workers), _sp_array(conf_dict),
                                                   _puct_array(conf_dict['puct_config']),
                                                   _puct_array(conf_dict['run_to_end_puct_config'])))
        ct = self.c_transformer
        shape = (ct.num_channels, -1, -1)
        for k in range(int(L.gz_supervisor_info(self._h, 0))):
            eng = PuctEngine.from_handle(L.gz_supervisor_engine(self._h, k), self.device, self.n_bases, self.role_sizes,
                                         (ct.num_channels, ct.channel_size, 1), plane_dtype=torch.float32, sm=self.sm)
            m = _Manager(eng, L.gz_supervisor_stream(self._h, k))
            if self.start_states_fn is not None:
                with m._ctx():
                    eng.set_start_states(self.start_states_fn(eng, k))
                m.stream.synchronize()
            self.managers.append(m)
        self.engine = self.managers[0].engine
        self.total_size = int(L.gz_supervisor_info(self._h, 2))
        self._views = {}

    # -- zero-copy path (inline mode): CUDA tensors in, CUDA tensors out, on the caller's stream
    def poll_device(self, policy=None, value=None, sub_iters=1):
        return self.engine.advance(policy, value, sub_iters)

    # -- reference-shaped numpy path
    def poll(self, n_predictions, arrays):
        """Returns a 1-D float32 numpy array VIEWING engine-owned pinned host memory (valid until the
        next poll that returns the same manager's batch, as in the reference: common.cpp:204-209), or
        None when nothing needs evaluating."""
        if not self.managers:
            raise GGPZeroInterfaceError('start_self_play() has not been called')
        if n_predictions:
            _check_predictions(arrays, n_predictions, self.role_sizes + [self.engine.n_rewards])
            for i, a in enumerate(arrays):
                self._ptrs[i] = a.ctypes.data
        n = self._L.gz_supervisor_poll(self._h, int(n_predictions), self._ptrs if n_predictions else None,
                                       len(self._ptrs) if n_predictions else 0, ctypes.byref(self._planes_ptr))
        if n < 0:
            _lib.check(n)
        if n == 0:
            return None
        addr = ctypes.addressof(self._planes_ptr.contents)
        view = self._views.get(addr)
        if view is None:
            view = self._views[addr] = np.ctypeslib.as_array(self._planes_ptr, shape=(self.batch_size * self.total_size,))
        return view[:n * self.total_size]

    def totals(self):
        from .engine import TOTAL_KEYS
        out = (ctypes.c_double * len(TOTAL_KEYS))()
        _lib.check(self._L.gz_supervisor_totals(self._h, out))
        return dict(zip(TOTAL_KEYS, [float(x) for x in out]))

    def check_health(self):
        """Raises when game slots were lost (halted) or finished samples were dropped (fetch_samples called too
        rarely for the ready buffer); the reference's sample vector is unbounded and its games never stop."""
        t = self.totals()
        if t['halted'] > 0 or t['dropped_samples'] > 0:
            raise GGPZeroInterfaceError('device self-play degraded: %d game slots halted (first error %d), %d samples '
                                        'dropped' % (t['halted'], t['first_error'], t['dropped_samples']))
        return t

    def fetch_samples(self):
        """``None`` or a list of dicts with the keys of ``supervisor_impl.cpp:75-118`` (== ``datadesc.Sample``)."""
        e = self.engine
        cap = e.ready_cap * len(self.managers)
        buf = np.zeros((cap, e.srec), dtype=np.int32)
        man = np.zeros(cap, dtype=np.int32)
        n = self._L.gz_supervisor_fetch_samples(self._h, buf.ctypes.data, cap, man.ctypes.data)
        if n < 0:
            _lib.check(int(n))
        res = []
        for k in range(len(self.managers)):
            recs = buf[:n][man[:n] == k]
            if len(recs):
                ident = self.identifier if len(self.managers) == 1 else '%s_m%d' % (self.identifier, k)
                res.extend(e.records_to_samples(recs, ident))
        return res or None

    def add_unique_state(self, bs):
        """The reference's Python wrapper no-ops this call as well (``cppinterface.py:218-219``)."""

    def clear_unique_states(self):
        _lib.check(self._L.gz_supervisor_clear_unique_states(self._h))


# ------------------------------------------------------------------------------------------------
class Player(object):
    """``ggpzero_interface.Player(sm, transformer, puct_conf_dict)``
    (``src/cpp/pyobjects/player_impl.cpp:8-129,183-201``, ``src/cpp/player.cpp``): a ctypes shell over the
    native ``gz_player`` (``csrc/gz_host.cu``) — the device engine in player mode (one tree).

    As in the reference, ``player_apply_move`` / ``player_move`` only *enqueue* work
    (``player.cpp:59-103``); it is carried out by the following ``poll`` calls, which hand out the
    planes of the leaf to evaluate and return ``None`` once the queue has drained. ``end_time`` and the
    configuration's ``think_time`` are honoured (the host watches the clock and tells the device search to stop
    at its next playout boundary, ``evaluator.cpp:794,820-826``).

    ``PuctConfig.batch_size`` > 1 (the reference's default is 32) runs the reference's intra-tree workers: a poll
    hands out up to ``batch_size`` leaves of the one tree and expects as many predictions back
    (``evaluator.cpp:722-742,1050-1079``, ``scheduler.cpp:131-203``).

    Differences from the reference: ``player_balance_moves`` needs ``batch_size`` > 1 (it runs on the
    search tree's workers); with ``batch_size`` 1 it is accepted and does nothing."""

    def __init__(self, sm, c_transformer, puct_conf_dict, device=0, seed=0x5EED0001, node_cap=1 << 16,
                 child_cap=None):
        from .engine import PuctEngine, _puct_array
        from .defs import confs
        import attr
        import torch
        if not isinstance(puct_conf_dict, dict):
            raise GGPZeroInterfaceError('puct config must be a dict (attr.asdict(PUCTEvaluatorConfig))')
        self._L = _lib.lib()
        _declare_host(self._L)
        L = self._L
        self.sm, self._pn = _resolve_sm(sm, device)
        self.conf = dict(attr.asdict(confs.PUCTEvaluatorConfig()))
        self.conf.update(puct_conf_dict)
        self.c_transformer = c_transformer
        self.n_bases = int(L.gz_propnet_info(self._pn, _lib.INFO_N_BASES))
        n_roles = int(L.gz_propnet_info(self._pn, _lib.INFO_N_ROLES))
        self.role_sizes = [int(L.gz_propnet_info(self._pn, _lib.INFO_ROLE_SIZE0 + r)) for r in range(n_roles)]
        c_transformer.finalize(self.n_bases, device)
        h = ctypes.c_void_p()
        _lib.check(L.gz_player_create(self._pn, c_transformer._h, _puct_array(self.conf), device, seed, node_cap,
                                      child_cap or 0, ctypes.byref(h)))
        self._h = h
        self.total_size = int(L.gz_player_info(h, 0))
        self.state_words = int(L.gz_player_info(h, 3))
        self.max_actions = int(L.gz_player_info(h, 2))
        self.rows_per_poll = int(L.gz_player_info(h, 4))      # PuctConfig.batch_size (most leaves one poll hands out)
        self.engine = PuctEngine.from_handle(L.gz_player_engine(h), device, self.n_bases, self.role_sizes,
                                             (c_transformer.num_channels, c_transformer.channel_size, 1),
                                             mode=PuctEngine.MODE_PLAYER, plane_dtype=torch.float32, sm=self.sm)
        self._ptrs = (ctypes.c_void_p * (n_roles + 1))()
        self._planes_ptr = ctypes.POINTER(ctypes.c_float)()
        self._planes_np = None

    def __del__(self):
        h = getattr(self, '_h', None)
        if h:
            self.engine = None
            self._L.gz_player_destroy(h)
            self._h = None

    # ---- reference method table
    def player_reset(self, game_depth):
        _lib.check(self._L.gz_player_reset(self._h, int(game_depth)))

    def player_update_config(self, think_time, converged_visits, verbose):
        self.conf['think_time'] = float(think_time)
        self.conf['converged_visits'] = int(converged_visits)
        self.conf['verbose'] = bool(verbose)
        _lib.check(self._L.gz_player_update_config(self._h, float(think_time), int(converged_visits), int(bool(verbose))))

    def player_apply_move(self, joint_move):
        if hasattr(joint_move, 'get') and hasattr(joint_move, 'get_count'):
            moves = [int(joint_move.get(i)) for i in range(joint_move.get_count())]
        else:
            moves = [int(x) for x in joint_move]
        arr = (ctypes.c_int * len(moves))(*moves)
        _lib.check(self._L.gz_player_apply_move(self._h, arr, len(moves)))

    def player_move(self, basestate, evaluations, end_time):
        from . import tables
        words = None
        if basestate is not None:
            if hasattr(basestate, 'to_list'):
                words = tables.pack_state_bits(np.asarray(basestate.to_list(), dtype=np.uint8), self.state_words)
            else:
                words = np.ascontiguousarray(basestate, dtype=np.uint32)
        _lib.check(self._L.gz_player_move(self._h, words.ctypes.data if words is not None else None, int(evaluations),
                                          float(end_time)))

    def player_get_move(self, lead_role_index):
        """-> (legal index of ``lead_role_index``'s move, the chosen child's score FOR THAT ROLE or -1, node count)
        (``player.cpp:105-119``)."""
        mv, sc, nodes = ctypes.c_int(), ctypes.c_float(), ctypes.c_int()
        _lib.check(self._L.gz_player_get_move(self._h, int(lead_role_index), ctypes.byref(mv), ctypes.byref(sc),
                                              ctypes.byref(nodes)))
        return (mv.value, float(sc.value), nodes.value)

    def player_balance_moves(self, max_count):
        """Accepted for API parity; the device engine does not rebalance first moves
        (``src/cpp/puct/evaluator.cpp:282-339`` is an analysis aid of the tournament tooling)."""
        _lib.check(self._L.gz_player_balance_moves(self._h, int(max_count)))

    def player_tree_debug(self, max_count):
        """-> tuple of (score, lead_role_index, move_index, variation) for the root's expanded children by traversals,
        ``variation`` = the principal line below each as ((lead_role_index, move_index), ..): at most 10 plies, through
        nodes with at least 100 visits (``src/cpp/puct/node.cpp:376-441``, ``player_impl.cpp:67-100``)."""
        n = max(0, int(max_count))
        m = max(n, 1)
        sc = np.zeros(m, dtype=np.float32)
        ld, mv, vl = np.zeros(m, dtype=np.int32), np.zeros(m, dtype=np.int32), np.zeros(m, dtype=np.int32)
        vp = np.zeros((m, 10, 2), dtype=np.int32)
        k = self._L.gz_player_tree_debug_lines(self._h, n, sc.ctypes.data, ld.ctypes.data, mv.ctypes.data,
                                               vl.ctypes.data, vp.ctypes.data)
        if k < 0:
            _lib.check(k)
        return tuple((float(sc[i]), int(ld[i]), int(mv[i]),
                      tuple((int(vp[i, j, 0]), int(vp[i, j, 1])) for j in range(int(vl[i])))) for i in range(k))

    def last_search(self):
        """Additive: the root's children after the last finished search (dict as ``PuctEngine.snapshot``), or None."""
        mc = self.max_actions
        oi = np.zeros(16, dtype=np.int32)
        of = np.zeros(8, dtype=np.float32)
        mv, tr = np.zeros(mc, dtype=np.int32), np.zeros(mc, dtype=np.int32)
        pr, sc, nx = np.zeros(mc, dtype=np.float32), np.zeros(mc, dtype=np.float32), np.zeros(mc, dtype=np.float32)
        n = self._L.gz_player_last_search(self._h, oi.ctypes.data, of.ctypes.data, mv.ctypes.data, tr.ctypes.data,
                                          pr.ctypes.data, sc.ctypes.data, nx.ctypes.data, mc)
        if n <= 0:
            return None
        R = len(self.role_sizes)
        return dict(choice=int(oi[3]), nodes=int(oi[4]), root_visits=int(oi[6]), lead_role=int(oi[7]),
                    other_moves=[int(x) for x in oi[9:9 + R]], root_score=float(of[0]),
                    choice_scores=[float(x) for x in of[1:1 + R]], moves=mv[:n].copy(), traversals=tr[:n].copy(),
                    policy_prob=pr[:n].copy(), scores=sc[:n].copy(), next_prob=nx[:n].copy())

    # ---- the poll contract (src/cpp/player.cpp:155-173, src/cpp/pyobjects/common.cpp:160-218)
    def poll(self, n_predictions, arrays):
        if n_predictions:
            _check_predictions(arrays, n_predictions, self.role_sizes + [self.engine.n_rewards])
            for i, a in enumerate(arrays):
                self._ptrs[i] = a.ctypes.data
        n = self._L.gz_player_poll(self._h, int(n_predictions), self._ptrs if n_predictions else None,
                                   len(self._ptrs) if n_predictions else 0, ctypes.byref(self._planes_ptr))
        if n < 0:
            _lib.check(n)
        if n == 0:
            return None
        if self._planes_np is None:
            self._planes_np = np.ctypeslib.as_array(self._planes_ptr, shape=(self.rows_per_poll * self.total_size,))
        return self._planes_np[:n * self.total_size]        # n leaves, flat (PuctConfig.batch_size at most)
