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
class Supervisor(object):
    """``ggpzero_interface.Supervisor(sm, transformer, batch_size, identifier)``
    (``src/cpp/pyobjects/supervisor_impl.cpp:215-237``) on the device engine.

    ``poll`` keeps the reference's numpy contract (``src/cpp/pyobjects/common.cpp:160-218``): it takes
    the predictions for the rows handed out by the previous poll and returns a 1-D float32 array of
    ``n_req * total_size`` planes (or None when there is nothing to evaluate). ``poll_device`` is the
    additive zero-copy variant on CUDA tensors."""

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
        self._rows = None
        self._primed = False

    def set_num_workers(self, n):
        """Not in the reference's method table although its Python wrapper calls it
        (``cppinterface.py:197-198``); accepted and ignored — the device engine has no worker threads."""

    def start_self_play(self, num_workers, conf_dict):
        from .engine import PuctEngine
        import torch
        self.engine = PuctEngine(self.game, conf_dict, n_games=self.batch_size, mode=0, device=self.device,
                                 seed=self.seed, node_cap=self.node_cap, child_cap=self.child_cap,
                                 c_transformer=self.c_transformer, plane_dtype=torch.float32,
                                 sm=self.sm if not isinstance(self.sm, str) else None)

    # -- zero-copy path
    def poll_device(self, policy=None, value=None, sub_iters=1):
        return self.engine.advance(policy, value, sub_iters)

    # -- reference-shaped numpy path
    def poll(self, n_predictions, arrays):
        """Returns a 1-D float32 numpy array VIEWING engine-owned pinned host memory (valid until the
        next poll, as in the reference: common.cpp:204-209), or None when nothing needs evaluating."""
        import torch
        e = self.engine
        if e is None:
            raise GGPZeroInterfaceError('start_self_play() has not been called')
        if not (0 <= n_predictions <= self.batch_size):
            raise GGPZeroInterfaceError('n_predictions out of range')
        if self._rows is None:
            if n_predictions != 0:
                raise GGPZeroInterfaceError('first poll must pass n_predictions == 0')
            self._host_planes = torch.empty((e.n_games, e.total_size), dtype=torch.float32).pin_memory()
            self._host_count = torch.zeros(1, dtype=torch.int64).pin_memory()
            self._stage_pol = torch.empty((e.n_games, e.policy_stride), dtype=torch.float32).pin_memory()
            self._stage_val = torch.empty((e.n_games, e.n_rewards), dtype=torch.float32).pin_memory()
            e.advance(None, None)
        else:
            if n_predictions != len(self._rows):
                raise GGPZeroInterfaceError('expected %d predictions, got %d' % (len(self._rows), n_predictions))
            if n_predictions:
                R = len(e.role_sizes)
                if len(arrays) != R + 1:
                    raise GGPZeroInterfaceError('expected %d policy arrays and one value array' % R)
                off = 0
                sp = self._stage_pol.numpy()
                for r in range(R):
                    a = arrays[r]
                    if not isinstance(a, np.ndarray) or a.dtype != np.float32 or not a.flags['C_CONTIGUOUS']:
                        raise GGPZeroInterfaceError('predictions must be C-contiguous float32 numpy arrays')
                    sp[:n_predictions, off:off + e.role_sizes[r]] = a.reshape(n_predictions, e.role_sizes[r])
                    off += e.role_sizes[r]
                self._stage_val.numpy()[:n_predictions] = arrays[R].reshape(n_predictions, e.n_rewards)
                rows = self._rows_dev
                e.policy.index_copy_(0, rows, self._stage_pol[:n_predictions].to(e.device, non_blocking=True))
                e.value.index_copy_(0, rows, self._stage_val[:n_predictions].to(e.device, non_blocking=True))
            e.advance(e.policy, e.value)
        need = self._need_eval()
        rows = torch.nonzero(need, as_tuple=False).flatten()        # device sync: the request count
        n_req = int(rows.numel())
        self._rows_dev = rows
        self._rows = rows                                           # only its length is used on the host
        if n_req == 0:
            return None
        gathered = e.planes.reshape(e.n_games, -1).index_select(0, rows).to(torch.float32)
        self._host_planes[:n_req].copy_(gathered, non_blocking=True)
        torch.cuda.current_stream(e.device).synchronize()
        return self._host_planes.numpy()[:n_req].reshape(-1)

    def _need_eval(self):
        import ctypes
        import torch
        e = self.engine
        if not hasattr(self, '_need_buf'):
            self._need_buf = torch.zeros(e.n_games, dtype=torch.uint8, device=e.device)
            e._L.gz_puct_copy_need_eval.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib.check(e._L.gz_puct_copy_need_eval(e._h, self._need_buf.data_ptr(), e._stream()))
        return self._need_buf

    def fetch_samples(self):
        recs = self.engine.fetch_sample_records()
        if len(recs) == 0:
            return None
        return self.engine.records_to_samples(recs, self.identifier)

    def add_unique_state(self, bs):
        """The reference's Python wrapper no-ops this call as well (``cppinterface.py:218-219``)."""

    def clear_unique_states(self):
        self.engine.clear_unique_states()
