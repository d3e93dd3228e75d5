"""Batched device state machine: the host-side mirror of the propnet entry points of
include/gz_b200.h. Device memory and streams come from torch; compute is libgz_b200's kernels.

Array conventions (torch tensors on the CUDA device):
    states   int32 [n, state_words]   packed bases (bit b -> word b>>5, bit b&31)
    legal    int32 [n, legal_words]   per role ceil(A_r/32) words at legal_word_off[r]
    moves    int16 [n, n_roles]
    terminal uint8 [n]       goals uint8 [n, n_roles]   (255 = none / not evaluated)

Reference counterpart: ggplib's StateMachineInterface as driven by the reference's C++ engine
(src/cpp/puct/node.cpp:153-221, src/cpp/puct/evaluator.cpp:216-239), batched over games.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import devprog, tables
from . import lib as _lib


class DeviceStateMachine:
    def __init__(self, game_or_propnet, device: int = 0, goals_mode: int = _lib.GOALS_ALWAYS):
        if not torch.cuda.is_available():
            raise _lib.GzError('no CUDA device: the propnet state machine has no CPU fallback')
        if isinstance(game_or_propnet, str):
            self.game = game_or_propnet
            self.propnet = tables.propnet_by_name(game_or_propnet)
        else:
            self.game = None
            self.propnet = game_or_propnet
        self.device = torch.device('cuda', device)
        self.goals_mode = goals_mode
        self._L = _lib.lib()
        blob = np.ascontiguousarray(devprog.pack_device_program(self.propnet), dtype=np.int32)
        h = ctypes.c_void_p()
        _lib.check(self._L.gz_propnet_create(blob.ctypes.data, len(blob), device, ctypes.byref(h)))
        # size of the three gate programs (next-state cone A, legal/terminal cone LT, goal cone GL): the work a
        # transition costs whatever the bytes it moves (bench.py reports the issue-rate bound from these)
        self.program_stats = {
            'gates': [int(blob[devprog.D_N_GATES_A]), int(blob[devprog.D_N_GATES_LT]), int(blob[devprog.D_N_GATES_GL])],
            'literals': [int(blob[devprog.D_N_LITS_A]), int(blob[devprog.D_N_LITS_LT]), int(blob[devprog.D_N_LITS_GL])],
            'chunks': [int(blob[devprog.D_CHUNKS_A]), int(blob[devprog.D_CHUNKS_LT]), int(blob[devprog.D_CHUNKS_GL])],
            'wide': bool(blob[devprog.D_WIDE]),
        }
        self._h = h
        info = lambda w: int(self._L.gz_propnet_info(self._h, w))  # noqa: E731
        self.n_bases = info(_lib.INFO_N_BASES)
        self.n_roles = info(_lib.INFO_N_ROLES)
        self.state_words = info(_lib.INFO_STATE_WORDS)
        self.legal_words = info(_lib.INFO_LEGAL_WORDS)
        self.role_sizes = [info(_lib.INFO_ROLE_SIZE0 + r) for r in range(self.n_roles)]
        self.legal_word_off = [info(_lib.INFO_LEGAL_WORD_OFF0 + r) for r in range(self.n_roles + 1)]
        self.launch_info = {
            'val_words': info(_lib.INFO_VAL_WORDS), 'warps_per_cta': info(_lib.INFO_WARPS_PER_CTA),
            'smem_bytes': info(_lib.INFO_SMEM_BYTES), 'grid': info(_lib.INFO_GRID),
            'staged_mask': info(_lib.INFO_STAGED_MASK),
        }
        self.init_words = tables.pack_state_bits(self.propnet.init_state, self.state_words)
        self.specialised = bool(info(10))      # a generated thread-sliced kernel exists for this game
        self.launch_info['specialised'] = self.specialised

    @property
    def sm_ptr(self) -> int:
        """The state machine as a pointer cast to an int — what the reference hands to its extension module
        (``cppinterface.py:12-30``); here the ``gz_propnet*`` of the device program."""
        return int(self._h.value)

    def set_group_min_batch(self, n: int):
        """Batches of at least ``n`` states use the group kernel (1024 games per CTA / cluster, one tile per lane;
        default 1024; < 0 disables, 0 forces it for every batch)."""
        self._L.gz_propnet_set_group_min_batch.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        _lib.check(self._L.gz_propnet_set_group_min_batch(self._h, int(n)))

    def set_deferred_goals_min_batch(self, n: int):
        """GOALS_TERMINAL_ONLY on a game with a deep goal cone: batches of at least ``n`` states evaluate the cone in a
        compacted second pass over the terminal states only (default 4096; < 0 disables)."""
        self._L.gz_propnet_set_deferred_goals_min_batch.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        _lib.check(self._L.gz_propnet_set_deferred_goals_min_batch(self._h, int(n)))

    def set_spec_min_batch(self, n: int):
        """Batches of at least ``n`` states use the game-specialised kernel (default 2^18; < 0 disables)."""
        self._L.gz_propnet_set_spec_min_batch.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        _lib.check(self._L.gz_propnet_set_spec_min_batch(self._h, int(n)))

    def __del__(self):
        h = getattr(self, '_h', None)
        if h:
            self._L.gz_propnet_destroy(h)
            self._h = None

    # ---- helpers
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def initial_states(self, n: int) -> torch.Tensor:
        w = torch.from_numpy(self.init_words.view(np.int32).copy()).to(self.device)
        return w.unsqueeze(0).repeat(n, 1).contiguous()

    def alloc_views(self, n: int):
        legal = torch.zeros((n, self.legal_words), dtype=torch.int32, device=self.device)
        terminal = torch.zeros((n,), dtype=torch.uint8, device=self.device)
        goals = torch.zeros((n, self.n_roles), dtype=torch.uint8, device=self.device)
        return legal, terminal, goals

    def _chk_states(self, t, n=None):
        assert t.dtype == torch.int32 and t.is_cuda and t.is_contiguous()
        assert t.dim() == 2 and t.shape[1] == self.state_words
        if n is not None:
            assert t.shape[0] == n

    # ---- batched entry points
    def eval(self, states, out=None):
        self._chk_states(states)
        n = states.shape[0]
        legal, terminal, goals = out if out is not None else self.alloc_views(n)
        _lib.check(self._L.gz_sm_eval(self._h, n, states.data_ptr(), legal.data_ptr(),
                                      terminal.data_ptr(), goals.data_ptr(), self.goals_mode,
                                      self._stream()))
        return legal, terminal, goals

    def transition(self, states, moves, out=None):
        self._chk_states(states)
        n = states.shape[0]
        assert moves.dtype == torch.int16 and moves.shape == (n, self.n_roles) and moves.is_contiguous()
        if out is None:
            nxt = torch.empty_like(states)
            legal, terminal, goals = self.alloc_views(n)
        else:
            nxt, legal, terminal, goals = out
        _lib.check(self._L.gz_sm_transition(self._h, n, states.data_ptr(), moves.data_ptr(),
                                            nxt.data_ptr(), legal.data_ptr(), terminal.data_ptr(),
                                            goals.data_ptr(), self.goals_mode, self._stream()))
        return nxt, legal, terminal, goals

    def playout_step(self, states, legal, terminal, goals, seed: int, step: int, slot0: int = 0,
                     moves_out=None, counter=None):
        self._chk_states(states)
        n = states.shape[0]
        _lib.check(self._L.gz_sm_playout_step(
            self._h, slot0, n, seed, step, states.data_ptr(), legal.data_ptr(), terminal.data_ptr(),
            goals.data_ptr(), moves_out.data_ptr() if moves_out is not None else None,
            counter.data_ptr() if counter is not None else None, self.goals_mode, self._stream()))

    def transition_host(self, states: np.ndarray, moves: np.ndarray, out=None):
        """Host buffers in, host buffers out (H2D + kernel + D2H inside the call)."""
        n = states.shape[0]
        assert states.dtype == np.uint32 and states.shape == (n, self.state_words)
        assert moves.dtype == np.int16 and moves.shape == (n, self.n_roles)
        if out is None:
            nxt = np.empty_like(states)
            legal = np.empty((n, self.legal_words), dtype=np.uint32)
            terminal = np.empty(n, dtype=np.uint8)
            goals = np.empty((n, self.n_roles), dtype=np.uint8)
        else:
            nxt, legal, terminal, goals = out
        _lib.check(self._L.gz_sm_transition_host(self._h, n, states.ctypes.data, moves.ctypes.data,
                                                 nxt.ctypes.data, legal.ctypes.data,
                                                 terminal.ctypes.data, goals.ctypes.data,
                                                 self.goals_mode))
        return nxt, legal, terminal, goals
