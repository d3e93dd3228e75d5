"""Board symmetries of states and actions — mirror of
``src/ggpzero/util/symmetry.py`` (``Translator``, ``Prescription``, ``create_translator``), plus the tables
that let the device apply a symmetry to a whole batch of packed states with one bit-permutation kernel
(``gz_permute_state_bits``) and to dense policy rows with one gather.

Base strings are ``(true (cell 1 1 red))`` and action strings ``(does black (move 1 1))`` (lookup.GameModel);
coordinates are translated exactly as the reference does: optional ``reflect_vertical`` first, then
``rot_count`` anti-clockwise quarter turns.
"""
from __future__ import annotations

import numpy as np

from .gdl.parser import parse_sexprs


def reflect_vertical(x, y, x_cords, y_cords):
    return x_cords[len(x_cords) - x_cords.index(x) - 1], y


def reflect_horizontal(x, y, x_cords, y_cords):
    return x, y_cords[len(y_cords) - y_cords.index(y) - 1]


def rotate_90(x, y, x_cords, y_cords):
    """anti-clockwise (symmetry.py:14-20)"""
    assert len(x_cords) == len(y_cords)
    x_idx, y_idx = x_cords.index(x), y_cords.index(y)
    return x_cords[y_idx], y_cords[len(x_cords) - x_idx - 1]


def symbolize(txt, pos):
    """``(true (cell 1 1 red))`` , 1 -> ('cell', '1', '1', 'red'); a bare term becomes a 1-tuple."""
    t = parse_sexprs(txt)[0][pos]
    return tuple(t) if isinstance(t, (tuple, list)) else (t,)


class Translator(object):
    def __init__(self, game_info, x_cords, y_cords):
        self.game_info = game_info
        self.x_cords = list(x_cords)
        self.y_cords = list(y_cords)
        self.base_symbols = [symbolize(b, 1) for b in game_info.model.bases]
        self.base_root_term_indexes = {}
        self.base_root_term_to_mapping = {}
        self.action_list = [[symbolize(a, 2) for a in game_info.model.actions[ri]]
                            for ri in range(len(game_info.model.roles))]
        self.action_index = [{t: i for i, t in enumerate(lst)} for lst in self.action_list]
        self.action_root_term_indexes = {}
        self.skip_base_root_terms = set()
        self.skip_action_root_term = set()

    @staticmethod
    def _aslist(v):
        return list(v) if isinstance(v, (list, tuple)) else [v]

    def add_basetype(self, root_term, x_terms_idx, y_terms_idx):
        assert root_term not in self.base_root_term_indexes
        self.base_root_term_indexes[root_term] = self._aslist(x_terms_idx), self._aslist(y_terms_idx)
        mapping = self.base_root_term_to_mapping[root_term] = {}
        for indx, terms in enumerate(self.base_symbols):
            if terms[0] == root_term:
                mapping[terms[1:]] = indx

    def add_action_type(self, root_term, x_terms_idx, y_terms_idx):
        assert root_term not in self.action_root_term_indexes
        self.action_root_term_indexes[root_term] = self._aslist(x_terms_idx), self._aslist(y_terms_idx)

    def add_skip_base(self, root_term):
        self.skip_base_root_terms.add(root_term)

    def add_skip_action(self, root_term):
        self.skip_action_root_term.add(root_term)

    def translate_terms(self, terms, x_terms_idx, y_terms_idx, do_reflection, rot_count):
        assert len(x_terms_idx) == len(y_terms_idx)
        new_terms = list(terms)
        for xi, yi in zip(x_terms_idx, y_terms_idx):
            x, y = terms[xi], terms[yi]
            if do_reflection:
                x, y = reflect_vertical(x, y, self.x_cords, self.y_cords)
            for _ in range(rot_count):
                x, y = rotate_90(x, y, self.x_cords, self.y_cords)
            new_terms[xi], new_terms[yi] = x, y
        return tuple(new_terms)

    def translate_base_index(self, indx, do_reflection, rot_count):
        terms = self.base_symbols[indx]
        if terms[0] in self.skip_base_root_terms:
            return indx
        if terms[0] not in self.base_root_term_indexes:
            raise Exception("Not supported base %s" % str(terms))
        xi, yi = self.base_root_term_indexes[terms[0]]
        new_terms = self.translate_terms(terms, xi, yi, do_reflection, rot_count)
        return self.base_root_term_to_mapping[new_terms[0]][new_terms[1:]]

    def translate_basestate(self, basestate, do_reflection, rot_count):
        assert isinstance(basestate, (tuple, list)) and len(basestate) == len(self.base_symbols)
        new_basestate = [0] * len(basestate)
        for indx in range(len(basestate)):
            if basestate[indx]:
                new_basestate[self.translate_base_index(indx, do_reflection, rot_count)] = 1
        return new_basestate

    translate_basestate_faster = translate_basestate

    def translate_action(self, role_index, legal, do_reflection, rot_count):
        terms = self.action_list[role_index][legal]
        if terms[0] in self.skip_action_root_term:
            return legal
        xi, yi = self.action_root_term_indexes[terms[0]]
        new_terms = self.translate_terms(terms, xi, yi, do_reflection, rot_count)
        return self.action_index[role_index][new_terms]

    # ---- tables for the batched device path
    def base_permutation(self, do_reflection, rot_count):
        """dest[i] = index the base i moves to (a permutation of range(n_bases))."""
        n = len(self.base_symbols)
        dest = np.array([self.translate_base_index(i, do_reflection, rot_count) for i in range(n)], dtype=np.int32)
        assert len(set(dest.tolist())) == n, 'a symmetry must permute the bases'
        return dest

    def action_permutation(self, role_index, do_reflection, rot_count):
        n = len(self.action_list[role_index])
        dest = np.array([self.translate_action(role_index, a, do_reflection, rot_count) for a in range(n)],
                        dtype=np.int32)
        assert len(set(dest.tolist())) == n
        return dest


class Prescription(object):
    """symmetry.py:228-254: which (reflection, rotation count) pairs a game uses."""

    def __init__(self, desc):
        self.prescription = []
        assert not (desc.do_rotations_90 and desc.do_rotations_180)
        if desc.do_rotations_90 or desc.do_rotations_180:
            rotations = (0, 2) if desc.do_rotations_180 else (0, 1, 2, 3)
            self.prescription += [(False, r) for r in rotations]
            if desc.do_reflection:
                self.prescription += [(True, r) for r in rotations]
        elif desc.do_reflection:
            self.prescription += [(False, 0), (True, 0)]
        else:
            self.prescription += [(False, 0)]

    def __iter__(self):
        return iter(self.prescription)


def create_translator(game_info, game_desc, game_symmetries):
    t = Translator(game_info, game_desc.x_cords, game_desc.y_cords)
    for ab in game_symmetries.apply_bases:
        t.add_basetype(ab.base_term, ab.x_terms_idx, ab.y_terms_idx)
    for ac in game_symmetries.apply_actions:
        t.add_action_type(ac.base_term, ac.x_terms_idx, ac.y_terms_idx)
    for term in game_symmetries.skip_bases:
        t.add_skip_base(term)
    for term in game_symmetries.skip_actions:
        t.add_skip_action(term)
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
