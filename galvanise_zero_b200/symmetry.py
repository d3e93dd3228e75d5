"""Board symmetries of states and actions.

The public surface is the one ``src/ggpzero/util/symmetry.py`` offers (``Translator`` with ``add_basetype`` /
``add_action_type`` / ``add_skip_base`` / ``add_skip_action`` / ``translate_basestate`` / ``translate_action``,
``Prescription``, ``create_translator``); the inside is built the way the device wants it: a symmetry is a
**permutation table**. Every base and every action is parsed once into (root term, board-coordinate slots as integer
indices into the game's coordinate lists); the eight board symmetries are 2x2 integer maps on those indices
(``reflect``: column c -> n-1-c; one anti-clockwise quarter turn: (c, r) -> (r, n-1-c)), applied to whole arrays with
numpy; the resulting tables ``dest[i]`` are cached per (reflection, rotation count) and shared by the host methods
(an index lookup) and by ``DeviceSymmetry`` (one bit-permutation kernel over a batch of packed states, one gather
over dense policy rows).

Base strings are ``(true (cell 1 1 red))``, action strings ``(does black (move 1 1))`` (lookup.GameModel). A
symmetry is applied as the reference applies it: the optional vertical-axis reflection first, then ``rot_count``
anti-clockwise quarter turns (symmetry.py:4-20,100-118).
"""
from __future__ import annotations

import numpy as np

from .gdl.parser import parse_sexprs


def _terms_of(txt, pos):
    t = parse_sexprs(txt)[0][pos]
    return tuple(t) if isinstance(t, (tuple, list)) else (t,)


def _slots(v):
    return tuple(v) if isinstance(v, (list, tuple)) else (v,)


class _TermTable(object):
    """One family of ground terms (the bases, or one role's actions) and its symmetry permutations."""

    def __init__(self, terms, n_cols, n_rows, col_of, row_of):
        self.terms = terms                        # tuple of symbols per entry, root term first
        self.index = {t: i for i, t in enumerate(terms)}
        self.n_cols, self.n_rows = n_cols, n_rows
        self.col_of, self.row_of = col_of, row_of
        self.moving = {}                          # root term -> (column slots, row slots)
        self.fixed = set()                        # root terms a symmetry leaves alone
        self._perm = {}

    def permutation(self, do_reflection, rot_count):
        key = (bool(do_reflection), int(rot_count) % 4)
        if key in self._perm:
            return self._perm[key]
        if key[1] and self.n_cols != self.n_rows:
            raise ValueError('quarter turns need a square board')
        n = self.n_cols
        dest = np.arange(len(self.terms), dtype=np.int32)
        x_names = sorted(self.col_of, key=self.col_of.get)
        y_names = sorted(self.row_of, key=self.row_of.get)
        for root, (xs, ys) in self.moving.items():
            members = [i for i, t in enumerate(self.terms) if t[0] == root]
            if not members:
                continue
            cols = np.array([[self.col_of[self.terms[i][k]] for k in xs] for i in members], dtype=np.int64)
            rows = np.array([[self.row_of[self.terms[i][k]] for k in ys] for i in members], dtype=np.int64)
            if key[0]:
                cols = self.n_cols - 1 - cols
            for _ in range(key[1]):
                cols, rows = rows, n - 1 - cols
            for m, i in enumerate(members):
                t = list(self.terms[i])
                for k, (xk, yk) in enumerate(zip(xs, ys)):
                    t[xk] = x_names[cols[m, k]]
                    t[yk] = y_names[rows[m, k]]
                dest[i] = self.index[tuple(t)]
        unknown = sorted({t[0] for t in self.terms} - set(self.moving) - self.fixed)
        if unknown:
            raise Exception("Not supported term(s) %s" % ', '.join(unknown))
        if len(np.unique(dest)) != len(dest):
            raise AssertionError('a symmetry must permute the terms')
        self._perm[key] = dest
        return dest


class Translator(object):
    def __init__(self, game_info, x_cords, y_cords):
        self.game_info = game_info
        self.x_cords, self.y_cords = list(x_cords), list(y_cords)
        col_of = {c: i for i, c in enumerate(self.x_cords)}
        row_of = {c: i for i, c in enumerate(self.y_cords)}
        model = game_info.model
        self._bases = _TermTable([_terms_of(b, 1) for b in model.bases], len(col_of), len(row_of), col_of, row_of)
        self._actions = [_TermTable([_terms_of(a, 2) for a in model.actions[ri]], len(col_of), len(row_of), col_of, row_of)
                         for ri in range(len(model.roles))]

    # ---- registration (what gamedesc.GameSymmetries prescribes)
    def add_basetype(self, root_term, x_terms_idx, y_terms_idx):
        assert root_term not in self._bases.moving
        xs, ys = _slots(x_terms_idx), _slots(y_terms_idx)
        assert len(xs) == len(ys)
        self._bases.moving[root_term] = (xs, ys)
        self._bases._perm.clear()

    def add_action_type(self, root_term, x_terms_idx, y_terms_idx):
        xs, ys = _slots(x_terms_idx), _slots(y_terms_idx)
        assert len(xs) == len(ys)
        for t in self._actions:
            assert root_term not in t.moving
            t.moving[root_term] = (xs, ys)
            t._perm.clear()

    def add_skip_base(self, root_term):
        self._bases.fixed.add(root_term)

    def add_skip_action(self, root_term):
        for t in self._actions:
            t.fixed.add(root_term)

    # ---- the tables
    def base_permutation(self, do_reflection, rot_count):
        """dest[i] = index base i moves to (a permutation of range(n_bases))."""
        return self._bases.permutation(do_reflection, rot_count)

    def action_permutation(self, role_index, do_reflection, rot_count):
        return self._actions[role_index].permutation(do_reflection, rot_count)

    @property
    def action_list(self):
        return [t.terms for t in self._actions]

    # ---- per-item queries are table look-ups
    def translate_base_index(self, indx, do_reflection, rot_count):
        return int(self.base_permutation(do_reflection, rot_count)[indx])

    def translate_basestate(self, basestate, do_reflection, rot_count):
        dest = self.base_permutation(do_reflection, rot_count)
        assert len(basestate) == len(dest)
        out = np.zeros(len(dest), dtype=np.int64)
        out[dest] = np.asarray(basestate, dtype=np.int64) != 0
        return out.tolist()

    translate_basestate_faster = translate_basestate

    def translate_action(self, role_index, legal, do_reflection, rot_count):
        return int(self.action_permutation(role_index, do_reflection, rot_count)[legal])


class Prescription(object):
    """Which (reflection, rotation count) pairs a game's description asks for (symmetry.py:228-254); the identity
    comes first, reflected variants after the plain ones."""

    def __init__(self, desc):
        assert not (desc.do_rotations_90 and desc.do_rotations_180)
        turns = (0, 1, 2, 3) if desc.do_rotations_90 else (0, 2) if desc.do_rotations_180 else (0,)
        mirrors = (False, True) if desc.do_reflection else (False,)
        self.prescription = [(m, r) for m in mirrors for r in turns]

    def __iter__(self):
        return iter(self.prescription)


def create_translator(game_info, game_desc, game_symmetries):
    t = Translator(game_info, game_desc.x_cords, game_desc.y_cords)
    for rule in game_symmetries.apply_bases:
        t.add_basetype(rule.base_term, rule.x_terms_idx, rule.y_terms_idx)
    for rule in game_symmetries.apply_actions:
        t.add_action_type(rule.base_term, rule.x_terms_idx, rule.y_terms_idx)
    for root in game_symmetries.skip_bases:
        t.add_skip_base(root)
    for root in game_symmetries.skip_actions:
        t.add_skip_action(root)
    return t


class DeviceSymmetry(object):
    """Applies one (reflection, rotation) to batches that live on the device: packed state rows through the
    bit-permutation kernel, dense policy rows through a gather."""

    def __init__(self, translator, do_reflection, rot_count, device=0):
        import torch
        self.device = torch.device('cuda', device)
        dest = translator.base_permutation(do_reflection, rot_count)
        src = np.empty_like(dest)
        src[dest] = np.arange(len(dest), dtype=np.int32)        # new bit j comes from old bit src[j]
        self.n_bases = len(dest)
        self.src = torch.from_numpy(src).to(self.device)
        self.action_src = []
        for ri in range(len(translator.action_list)):
            d = translator.action_permutation(ri, do_reflection, rot_count)
            s = np.empty_like(d)
            s[d] = np.arange(len(d), dtype=np.int32)
            self.action_src.append(torch.from_numpy(s.astype(np.int64)).to(self.device))

    def states(self, states, out=None):
        """int32 [n, state_words] packed states -> the symmetric states (same layout)."""
        import ctypes
        import torch
        from . import lib as _lib
        L = _lib.lib()
        assert states.is_cuda and states.dtype == torch.int32 and states.is_contiguous()
        out = torch.empty_like(states) if out is None else out
        L.gz_permute_state_bits.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        stream = ctypes.c_void_p(torch.cuda.current_stream(states.device).cuda_stream)
        _lib.check(L.gz_permute_state_bits(states.shape[0], states.shape[1], self.n_bases, states.data_ptr(),
                                           self.src.data_ptr(), out.data_ptr(), stream))
        return out

    def policy(self, role_index, policy):
        """dense policy rows [n, A_r]: new[:, dest[a]] = old[:, a]."""
        return policy.index_select(1, self.action_src[role_index])
