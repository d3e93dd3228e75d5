"""ORACLE (test infrastructure): ctypes binding of oracle/libpnref.so (propnet_ref.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(['make', '-s', '-C', HERE, 'libpnref.so'])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, 'libpnref.so')
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        vp, i32p, u32p, u8p, i16p = (ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32),
                                      ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint8),
                                      ctypes.POINTER(ctypes.c_int16))
        L.pnref_create.restype = vp
        L.pnref_create.argtypes = [i32p, ctypes.c_int64]
        L.pnref_destroy.argtypes = [vp]
        L.pnref_dupe.restype = vp
        L.pnref_dupe.argtypes = [vp]
        for name in ('pnref_n_bases', 'pnref_n_roles', 'pnref_state_words', 'pnref_legal_words',
                     'pnref_is_terminal'):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = ctypes.c_int
        L.pnref_role_size.argtypes = [vp, ctypes.c_int]
        L.pnref_initial_state.argtypes = [vp, u32p]
        L.pnref_update_bases.argtypes = [vp, u32p]
        L.pnref_legal_count.argtypes = [vp, ctypes.c_int]
        L.pnref_legal.argtypes = [vp, ctypes.c_int, ctypes.c_int]
        L.pnref_goal.argtypes = [vp, ctypes.c_int]
        L.pnref_next_state.argtypes = [vp, i32p, u32p]
        L.pnref_eval_batch.argtypes = [vp, ctypes.c_int64, u32p, u32p, u8p, u8p]
        L.pnref_transition_batch.argtypes = [vp, ctypes.c_int64, u32p, i16p, u32p, u32p, u8p, u8p]
        L.pnref_playout_step.restype = ctypes.c_int64
        L.pnref_playout_step.argtypes = [vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_uint64,
                                         ctypes.c_uint64, u32p, u32p, u8p, u8p, i16p]
        L.pnref_playout_bench.restype = ctypes.c_double
        L.pnref_playout_bench.argtypes = [i32p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_uint64, ctypes.c_int,
                                          ctypes.POINTER(ctypes.c_int64),
                                          ctypes.POINTER(ctypes.c_uint64)]
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


class RefStateMachine:
    """Scalar CPU propnet state machine with the ggplib call protocol."""

    def __init__(self, blob: np.ndarray):
        self.blob = np.ascontiguousarray(blob, dtype=np.int32)
        self.L = lib()
        self.h = self.L.pnref_create(_p(self.blob, ctypes.c_int32), len(self.blob))
        if not self.h:
            raise ValueError('bad propnet blob')
        self.n_bases = self.L.pnref_n_bases(self.h)
        self.n_roles = self.L.pnref_n_roles(self.h)
        self.state_words = self.L.pnref_state_words(self.h)
        self.legal_words = self.L.pnref_legal_words(self.h)
        self.role_sizes = [self.L.pnref_role_size(self.h, r) for r in range(self.n_roles)]

    def __del__(self):
        if getattr(self, 'h', None):
            self.L.pnref_destroy(self.h)
            self.h = None

    def initial_state(self):
        out = np.zeros(self.state_words, dtype=np.uint32)
        self.L.pnref_initial_state(self.h, _p(out, ctypes.c_uint32))
        return out

    def update_bases(self, state):
        state = np.ascontiguousarray(state, dtype=np.uint32)
        self.L.pnref_update_bases(self.h, _p(state, ctypes.c_uint32))

    def legals(self, role):
        n = self.L.pnref_legal_count(self.h, role)
        return [self.L.pnref_legal(self.h, role, i) for i in range(n)]

    def is_terminal(self):
        return bool(self.L.pnref_is_terminal(self.h))

    def goal(self, role):
        return self.L.pnref_goal(self.h, role)

    def next_state(self, joint_move):
        jm = np.ascontiguousarray(joint_move, dtype=np.int32)
        out = np.zeros(self.state_words, dtype=np.uint32)
        self.L.pnref_next_state(self.h, _p(jm, ctypes.c_int32), _p(out, ctypes.c_uint32))
        return out

    # ---- batch
    def eval_batch(self, states):
        states = np.ascontiguousarray(states, dtype=np.uint32)
        n = states.shape[0]
        legal = np.zeros((n, self.legal_words), dtype=np.uint32)
        term = np.zeros(n, dtype=np.uint8)
        goals = np.zeros((n, self.n_roles), dtype=np.uint8)
        self.L.pnref_eval_batch(self.h, n, _p(states, ctypes.c_uint32), _p(legal, ctypes.c_uint32),
                                _p(term, ctypes.c_uint8), _p(goals, ctypes.c_uint8))
        return legal, term, goals

    def transition_batch(self, states, moves):
        states = np.ascontiguousarray(states, dtype=np.uint32)
        moves = np.ascontiguousarray(moves, dtype=np.int16)
        n = states.shape[0]
        nxt = np.zeros((n, self.state_words), dtype=np.uint32)
        legal = np.zeros((n, self.legal_words), dtype=np.uint32)
        term = np.zeros(n, dtype=np.uint8)
        goals = np.zeros((n, self.n_roles), dtype=np.uint8)
        self.L.pnref_transition_batch(self.h, n, _p(states, ctypes.c_uint32),
                                      _p(moves, ctypes.c_int16), _p(nxt, ctypes.c_uint32),
                                      _p(legal, ctypes.c_uint32), _p(term, ctypes.c_uint8),
                                      _p(goals, ctypes.c_uint8))
        return nxt, legal, term, goals

    def playout_step(self, slot0, seed, step, states, legal, term, goals, moves_out=None):
        n = states.shape[0]
        mp = _p(moves_out, ctypes.c_int16) if moves_out is not None else None
        return self.L.pnref_playout_step(self.h, slot0, n, seed, step, _p(states, ctypes.c_uint32),
                                         _p(legal, ctypes.c_uint32), _p(term, ctypes.c_uint8),
                                         _p(goals, ctypes.c_uint8), mp)


def playout_bench(blob, n_slots, n_steps, seed, n_threads):
    blob = np.ascontiguousarray(blob, dtype=np.int32)
    tr = ctypes.c_int64(0)
    ck = ctypes.c_uint64(0)
    secs = lib().pnref_playout_bench(_p(blob, ctypes.c_int32), len(blob), n_slots, n_steps, seed,
                                     n_threads, ctypes.byref(tr), ctypes.byref(ck))
    return secs, tr.value, ck.value


# ------------------------------------------------------------------------------------------------
class RefEncoder:
    """ctypes binding of oracle/encoder_ref.c (restatement of gdltransformer.cpp:14-52)."""

    def __init__(self, channel_size, channels_per_state, num_control_channels, num_prev_states):
        L = lib()
        vp, ci = ctypes.c_void_p, ctypes.c_int
        L.encref_create.restype = vp
        L.encref_create.argtypes = [ci, ci, ci, ci]
        L.encref_destroy.argtypes = [vp]
        L.encref_add_board_base.argtypes = [vp, ci, ci]
        L.encref_add_control_base.argtypes = [vp, ci, ci, ctypes.c_float]
        L.encref_total_size.argtypes = [vp]
        L.encref_batch.argtypes = [vp, ctypes.c_int64, ci, vp, vp, vp, ci, vp]
        self.L = L
        self.h = L.encref_create(channel_size, channels_per_state, num_control_channels, num_prev_states)

    @classmethod
    def from_c_transformer(cls, ct):
        e = cls(ct.channel_size, ct.channels_per_state, ct.num_control_channels, ct.num_prev_states)
        for b, off in ct.board_bases:
            e.L.encref_add_board_base(e.h, b, off)
        for b, c, v in ct.control_bases:
            e.L.encref_add_control_base(e.h, b, c, v)
        return e

    def __del__(self):
        if getattr(self, 'h', None):
            self.L.encref_destroy(self.h)
            self.h = None

    def total_size(self):
        return self.L.encref_total_size(self.h)

    def encode(self, states, prev=None, prev_count=None):
        states = np.ascontiguousarray(states, dtype=np.uint32)
        n, sw = states.shape
        out = np.empty((n, self.total_size()), dtype=np.float32)
        max_prev = 0
        pp = pc = None
        if prev is not None:
            prev = np.ascontiguousarray(prev, dtype=np.uint32)
            max_prev = prev.shape[1]
            pp = prev.ctypes.data
            if prev_count is None:
                prev_count = np.full(n, max_prev, dtype=np.uint8)
            prev_count = np.ascontiguousarray(prev_count, dtype=np.uint8)
            pc = prev_count.ctypes.data
        self.L.encref_batch(self.h, n, sw, states.ctypes.data, pp, pc, max_prev, out.ctypes.data)
        return out
