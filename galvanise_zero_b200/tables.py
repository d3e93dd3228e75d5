"""Propnet tables: game lookup by name and packing into the flat int32 blob of
``include/gz_propnet_format.h`` (the only thing that crosses the C-ABI for a game definition).

Stands in for ``ggplib.db.lookup.by_name(game)`` as used by the reference
(``src/ggpzero/nn/bases.py:26-29``, ``src/ggpzero/util/symmetry.py:47``).
"""
from __future__ import annotations

import os
from functools import lru_cache

import numpy as np

from .gdl.propnet import Propnet, load_propnet

GAMES_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'games')

# header indices (mirror include/gz_propnet_format.h)
MAGIC = 0x475A504E
VERSION = 3
HEADER_WORDS = 48
(H_MAGIC, H_VERSION, H_N_BASES, H_N_ROLES, H_N_SLOTS, H_N_GATES, H_TERMINAL_LIT, H_STATE_WORDS,
 H_LEGAL_WORDS, H_N_INPUTS, H_TOTAL_WORDS, H_OFF_ROLE_SIZES, H_OFF_GATE_OUT, H_OFF_LIT_OFF,
 H_OFF_LIT_DATA, H_OFF_NEXT_LIT, H_OFF_LEGAL_OFF, H_OFF_LEGAL_LIT, H_OFF_GOAL_OFF, H_OFF_GOAL_VAL,
 H_OFF_GOAL_LIT, H_OFF_INIT, H_OFF_PROG_V, H_LEN_PROG_V, H_OFF_PROG_T, H_LEN_PROG_T,
 H_OFF_PROG_TV, H_LEN_PROG_TV, H_OFF_PROG_LT, H_LEN_PROG_LT, H_OFF_PROG_GL, H_LEN_PROG_GL,
 H_OFF_LEGAL_WORD_OFF, H_OFF_PROG_TL, H_LEN_PROG_TL) = range(35)


def available_games():
    return sorted(f[:-4] for f in os.listdir(GAMES_DIR) if f.endswith('.npz'))


@lru_cache(maxsize=None)
def propnet_by_name(game: str) -> Propnet:
    path = os.path.join(GAMES_DIR, game + '.npz')
    if not os.path.exists(path):
        raise KeyError('unknown game %r (compiled tables: %s); compile a rulesheet with '
                       'tools/compile_games.py' % (game, ', '.join(available_games())))
    return load_propnet(path)


def pad4(n: int) -> int:
    return (n + 3) & ~3


def state_words(n_bases: int) -> int:
    return pad4((n_bases + 31) // 32)


def legal_layout(role_sizes):
    offs = [0]
    for a in role_sizes:
        offs.append(offs[-1] + (a + 31) // 32)
    return offs, pad4(offs[-1])


def pack_state_bits(bits: np.ndarray, n_words: int) -> np.ndarray:
    """uint8 0/1 array [..., n_bits] -> uint32 [..., n_words], LSB-first."""
    bits = np.asarray(bits, dtype=np.uint8)
    n = bits.shape[-1]
    padded = np.zeros(bits.shape[:-1] + (n_words * 32,), dtype=np.uint8)
    padded[..., :n] = bits
    by = np.packbits(padded, axis=-1, bitorder='little')
    return np.ascontiguousarray(by).view('<u4').reshape(bits.shape[:-1] + (n_words,))


def unpack_state_bits(words: np.ndarray, n_bits: int) -> np.ndarray:
    words = np.ascontiguousarray(np.asarray(words, dtype='<u4'))
    by = words.view(np.uint8).reshape(words.shape[:-1] + (words.shape[-1] * 4,))
    return np.unpackbits(by, axis=-1, bitorder='little')[..., :n_bits]


def pack_blob(pn: Propnet) -> np.ndarray:
    nr = len(pn.roles)
    role_sizes = pn.role_sizes
    sw = state_words(pn.n_bases)
    lwo, lw = legal_layout(role_sizes)
    sections = []
    hdr = np.zeros(HEADER_WORDS, dtype=np.int64)
    pos = HEADER_WORDS

    def add(arr, off_idx, len_idx=None):
        nonlocal pos
        arr = np.asarray(arr).astype(np.int64).ravel()
        hdr[off_idx] = pos
        if len_idx is not None:
            hdr[len_idx] = len(arr)
        sections.append(arr)
        pos += len(arr)

    lit_off = np.zeros(len(pn.gate_lits) + 1, dtype=np.int64)
    for i, ls in enumerate(pn.gate_lits):
        lit_off[i + 1] = lit_off[i] + len(ls)
    lit_data = np.concatenate(pn.gate_lits) if len(pn.gate_lits) else np.zeros(0, np.int32)
    legal_off = np.concatenate([[0], np.cumsum(role_sizes)])
    goal_off = np.concatenate([[0], np.cumsum([len(v) for v in pn.goal_values])])
    add(role_sizes, H_OFF_ROLE_SIZES)
    add(pn.gate_out, H_OFF_GATE_OUT)
    add(lit_off, H_OFF_LIT_OFF)
    add(lit_data, H_OFF_LIT_DATA)
    add(pn.next_lit, H_OFF_NEXT_LIT)
    add(legal_off, H_OFF_LEGAL_OFF)
    add(np.concatenate(pn.legal_lit), H_OFF_LEGAL_LIT)
    add(goal_off, H_OFF_GOAL_OFF)
    add(np.concatenate(pn.goal_values) if nr else [], H_OFF_GOAL_VAL)
    add(np.concatenate(pn.goal_lit) if nr else [], H_OFF_GOAL_LIT)
    add(pack_state_bits(pn.init_state, sw).astype(np.int64), H_OFF_INIT)
    add(pn.prog['V'], H_OFF_PROG_V, H_LEN_PROG_V)
    add(pn.prog['T'], H_OFF_PROG_T, H_LEN_PROG_T)
    add(pn.prog['TV'], H_OFF_PROG_TV, H_LEN_PROG_TV)
    add(pn.prog['LT'], H_OFF_PROG_LT, H_LEN_PROG_LT)
    add(pn.prog['GL'], H_OFF_PROG_GL, H_LEN_PROG_GL)
    add(lwo, H_OFF_LEGAL_WORD_OFF)
    add(pn.prog['TL'], H_OFF_PROG_TL, H_LEN_PROG_TL)
    hdr[H_MAGIC] = MAGIC
    hdr[H_VERSION] = VERSION
    hdr[H_N_BASES] = pn.n_bases
    hdr[H_N_ROLES] = nr
    hdr[H_N_SLOTS] = pn.n_slots
    hdr[H_N_GATES] = len(pn.gate_out)
    hdr[H_TERMINAL_LIT] = pn.terminal_lit
    hdr[H_STATE_WORDS] = sw
    hdr[H_LEGAL_WORDS] = lw
    hdr[H_N_INPUTS] = sum(role_sizes)
    hdr[H_TOTAL_WORDS] = pos
    blob = np.concatenate([hdr] + sections)
    # the packed init state holds uint32 bit patterns; keep them through the int32 view
    return (blob & 0xFFFFFFFF).astype(np.uint32).view(np.int32)


@lru_cache(maxsize=None)
def blob_by_name(game: str) -> np.ndarray:
    return pack_blob(propnet_by_name(game))
