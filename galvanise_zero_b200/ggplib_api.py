"""ggplib-shaped single-state API on top of the batched device state machine.

The reference's Python code drives ggplib's state machine one state at a time
(``src/test/nn/test_model.py:41-55``, ``src/ggpzero/player/puctplayer.py:71-79``,
``src/ggpzero/scripts/findbases.py:121-153``): ``update_bases``, ``get_legal_state(ri)`` ->
``get_count / get_legal(i)``, ``is_terminal``, ``get_goal_value(ri)``, ``next_state(jm, out)``,
``get_joint_move`` -> ``get / set``, ``new_base_state``, ``get_initial_state``, ``get_current_state``,
``reset``, ``get_roles``, ``legal_to_move``, ``dupe``; ``BaseState`` offers ``to_list / from_list /
to_string / from_string / assign / hash_code / get / set / ==``. Every call below is a batch-of-one
launch of the CUDA kernels (API parity for tests and match play — throughput paths use the batched
engine directly).
"""
from __future__ import annotations

import numpy as np
import torch

from . import tables
from .statemachine import DeviceStateMachine


class BaseState(object):
    def __init__(self, n_bases, words=None):
        self.len = n_bases
        self.n_words = tables.state_words(n_bases)
        self.words = np.zeros(self.n_words, dtype=np.uint32) if words is None else np.array(words, dtype=np.uint32)

    def get(self, i):
        return bool((self.words[i >> 5] >> np.uint32(i & 31)) & 1)

    def set(self, i, v):
        if v:
            self.words[i >> 5] |= np.uint32(1 << (i & 31))
        else:
            self.words[i >> 5] &= np.uint32(~(1 << (i & 31)) & 0xFFFFFFFF)

    def to_list(self):
        return [int(x) for x in tables.unpack_state_bits(self.words, self.len)]

    def from_list(self, lst):
        assert len(lst) == self.len
        self.words = tables.pack_state_bits(np.array(lst, dtype=np.uint8), self.n_words)

    def to_string(self):
        return self.words.tobytes()

    @property
    def ptr(self):
        """The state as the extension module takes it (``ggpzero_interface.Player.player_move(basestate_ptr, ..)``):
        the address of the packed uint32 words — the counterpart of ggplib's ``ffi.cast("intptr_t", bs)``
        (``cppinterface.py:12-30``). Valid while this object is alive and unmodified."""
        self.words = np.ascontiguousarray(self.words, dtype=np.uint32)
        return int(self.words.ctypes.data)

    def from_string(self, s):
        self.words = np.frombuffer(s, dtype=np.uint32).copy()

    def assign(self, other):
        self.words = other.words.copy()

    def hash_code(self):
        return hash(self.words.tobytes())

    def __eq__(self, other):
        return isinstance(other, BaseState) and np.array_equal(self.words, other.words)

    def __hash__(self):
        return self.hash_code()


class LegalState(object):
    def __init__(self, legals):
        self.legals = legals

    def get_count(self):
        return len(self.legals)

    def get_legal(self, i):
        return self.legals[i]

    def to_list(self):
        return list(self.legals)


class JointMove(object):
    def __init__(self, n_roles):
        self.moves = np.zeros(n_roles, dtype=np.int32)

    def get(self, ri):
        return int(self.moves[ri])

    def set(self, ri, v):
        self.moves[ri] = int(v)

    def get_count(self):
        return len(self.moves)

    @property
    def ptr(self):
        """The joint move as the extension module takes it (``player_apply_move(joint_move_ptr)``): the address of
        one int32 per role."""
        return int(self.moves.ctypes.data)


class StateMachine(object):
    def __init__(self, game_info, device=0, _dev=None):
        self.game_info = game_info
        self.game = game_info.game
        self.dev = _dev or DeviceStateMachine(self.game, device=device)
        self._h = self.dev._h                       # so that the engine can share the device program
        self.sm_ptr = self.dev.sm_ptr               # what ggpzero_interface.Supervisor / Player take
        self.role_sizes = self.dev.role_sizes
        self.n_roles = self.dev.n_roles
        self._joint_move = JointMove(self.n_roles)
        self._current = self.get_initial_state()
        self._views = None

    # -- construction helpers
    def dupe(self):
        sm = StateMachine(self.game_info, _dev=self.dev)
        sm._current.assign(self._current)
        sm._views = self._views
        return sm

    def new_base_state(self):
        return BaseState(self.dev.n_bases)

    def get_initial_state(self):
        return BaseState(self.dev.n_bases, self.dev.init_words)

    def get_current_state(self, bs=None):
        if bs is not None:
            bs.assign(self._current)
            return bs
        cur = self.new_base_state()
        cur.assign(self._current)
        return cur

    def get_roles(self):
        return list(self.game_info.model.roles)

    def get_role_count(self):
        return self.n_roles

    def get_joint_move(self):
        return JointMove(self.n_roles)

    def legal_to_move(self, ri, legal):
        return self.game_info.propnet.inputs[ri][legal]

    # -- the state-machine calls
    def _state_tensor(self, bs):
        return torch.from_numpy(bs.words.view(np.int32).reshape(1, -1).copy()).to(self.dev.device)

    def update_bases(self, bs):
        self._current.assign(bs)
        legal, term, goals = self.dev.eval(self._state_tensor(bs))
        self._views = (legal.cpu().numpy().view(np.uint32)[0], int(term.item()), goals.cpu().numpy()[0])

    def reset(self):
        self.update_bases(self.get_initial_state())

    def _need_views(self):
        if self._views is None:
            self.update_bases(self._current)
        return self._views

    def get_legal_state(self, ri):
        legal, _, _ = self._need_views()
        o0, o1 = self.dev.legal_word_off[ri], self.dev.legal_word_off[ri + 1]
        bits = tables.unpack_state_bits(legal[o0:o1], self.role_sizes[ri])
        return LegalState([int(i) for i in np.nonzero(bits)[0]])

    def is_terminal(self):
        return bool(self._need_views()[1])

    def get_goal_value(self, ri):
        g = int(self._need_views()[2][ri])
        return -1 if g == 255 else g

    def next_state(self, joint_move, out_bs):
        moves = torch.tensor([[joint_move.get(r) for r in range(self.n_roles)]], dtype=torch.int16,
                             device=self.dev.device)
        nxt, _, _, _ = self.dev.transition(self._state_tensor(self._current), moves)
        out_bs.words = nxt.cpu().numpy().view(np.uint32)[0].copy()
        return out_bs
