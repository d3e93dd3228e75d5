"""Lowering of a compiled propnet to the *device program* consumed by the CUDA kernels.

A device program is a stream of 32-gate chunks. One warp owns a tile of 32 games; the value of a
gate for those 32 games is one 32-bit word in the warp's value array (shared memory). In a chunk,
lane ``i`` evaluates gate ``i``:  ``val[out_i] = AND_k (val[lit_ik >> 1] ^ -(lit_ik & 1))`` with
``nl`` literal rows per chunk (uniform per chunk, padded with the TRUE literal), and a warp barrier
follows the last chunk of every level.

Stream layout (uint16 units), per chunk:  [nl | sync<<8, 0,  out[32],  lit rows: nl x 32]
Literal rows are stored pair-interleaved so that a lane fetches two literals with one 32-bit load:
row pair p holds for lane i the uint32  lit[2p][i] | lit[2p+1][i] << 16  (nl is even).

Two independent phase programs are emitted, each over its own slot space (leaves first):
  * phase A ("T"):  bases + inputs -> next-state literals
  * phase B ("V"):  bases -> legal / terminal (sub-stream LT) and goal (sub-stream GL) literals
Slots of temporaries are re-used once their last reader's level has completed, which keeps the
per-warp value array small (reversi: 21.8 K gates -> ~1 K live slots).

Reference counterpart: none — ggplib keeps one mutable propnet per state machine instance; this is
the batched, bit-sliced re-design of the work done by StateMachineInterface::updateBases/nextState
(called at src/cpp/puct/evaluator.cpp:218-219, src/cpp/puct/node.cpp:157).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import tables
from .gdl.propnet import Propnet

DEV_MAGIC = 0x475A4450  # 'GZDP'
DEV_VERSION = 3
DEV_HEADER_WORDS = 64
MAX_SLOT = 32767 - 32   # 15-bit slot index, minus the 32 dump slots

(D_MAGIC, D_VERSION, D_N_BASES, D_N_ROLES, D_N_INPUTS, D_STATE_WORDS, D_LEGAL_WORDS,
 D_TOTAL_WORDS, D_VAL_WORDS, D_TERMINAL_LIT,
 D_OFF_ROLE_SIZES, D_OFF_LEGAL_WORD_OFF, D_OFF_INIT,
 D_OFF_PROG_A, D_CHUNKS_A, D_HALFS_A,
 D_OFF_PROG_LT, D_CHUNKS_LT, D_HALFS_LT,
 D_OFF_PROG_GL, D_CHUNKS_GL, D_HALFS_GL,
 D_OFF_NEXT_LIT, D_OFF_LEGAL_LIT, D_OFF_GOAL_OFF, D_OFF_GOAL_VAL, D_OFF_GOAL_LIT,
 D_N_GATES_A, D_N_GATES_LT, D_N_GATES_GL, D_N_LITS_A, D_N_LITS_LT, D_N_LITS_GL,
 D_DUMP_SLOT, D_WIDE) = range(35)


class _Alloc:
    """Slot allocator with release at level boundaries."""

    def __init__(self, first_free: int):
        self.next = first_free
        self.free: List[int] = []
        self.high = first_free

    def get(self) -> int:
        if self.free:
            return self.free.pop()
        s = self.next
        self.next += 1
        self.high = max(self.high, self.next)
        return s

    def put(self, s: int):
        self.free.append(s)


def _pack_phase(pn: Propnet, parts: List[Sequence[int]], keep_lits: Sequence[int],
                n_leaves: int, wide: bool = False) -> Tuple[List[np.ndarray], Dict[int, int], int, List[Tuple[int, int]]]:
    """Allocate slots and emit chunk streams for consecutive sub-programs ``parts`` that share one
    slot space. ``keep_lits``: output literals (old slot space) that must stay readable at the end.
    Returns (streams per part, old_slot->new_slot map for kept slots, slots used, stats per part)."""
    all_gates: List[int] = [int(g) for p in parts for g in p]
    part_of = {}
    for pi, p in enumerate(parts):
        for g in p:
            part_of[int(g)] = pi
    out_to_gate = {int(pn.gate_out[g]): g for g in all_gates}
    # level numbering: parts run one after another, so offset each part's levels
    level: Dict[int, int] = {}
    base_level = 0
    seq_levels: List[List[int]] = []     # list of levels, each a list of gates (in execution order)
    part_level_ranges = []
    for p in parts:
        lv = _levelise_part(pn, [int(g) for g in p], out_to_gate, level, base_level)
        n_lv = (max(lv.values()) - base_level) if lv else 0
        start = len(seq_levels)
        for _ in range(n_lv):
            seq_levels.append([])
        for g in p:
            seq_levels[start + lv[int(g)] - base_level - 1].append(int(g))
        level.update(lv)
        part_level_ranges.append((start, start + n_lv))
        base_level += n_lv
    # last use (sequence level index) of every gate output
    keep_slots = set(int(l) >> 1 for l in keep_lits)
    last_use: Dict[int, int] = {}
    for li, gates in enumerate(seq_levels):
        for g in gates:
            for l in pn.gate_lits[g]:
                h = out_to_gate.get(int(l) >> 1)
                if h is not None:
                    last_use[h] = li
    release_at: Dict[int, List[int]] = {}
    for g, li in last_use.items():
        if int(pn.gate_out[g]) not in keep_slots:
            release_at.setdefault(li, []).append(g)
    alloc = _Alloc(n_leaves)
    new_slot: Dict[int, int] = {}        # gate -> new slot

    def map_lit(l: int) -> int:
        s = int(l) >> 1
        h = out_to_gate.get(s)
        if h is None:
            if s >= n_leaves:
                raise ValueError('literal refers to a gate outside the phase program')
            return int(l)
        return new_slot[h] * 2 | (int(l) & 1)

    streams: List[np.ndarray] = []
    stats = []
    pending_release: List[int] = []
    for pi, (ls, le) in enumerate(part_level_ranges):
        chunks: List[np.ndarray] = []
        n_g = n_l = 0
        for li in range(ls, le):
            gates = seq_levels[li]
            # a gate nobody reads and nobody keeps can still be emitted (dead code is removed upstream)
            gates = sorted(gates, key=lambda g: len(pn.gate_lits[g]))
            # slots released by levels < li are available now
            for s in pending_release:
                alloc.put(s)
            pending_release = []
            for g in gates:
                new_slot[g] = alloc.get()
            for c0 in range(0, len(gates), 32):
                cg = gates[c0:c0 + 32]
                nl = max(len(pn.gate_lits[g]) for g in cg)
                last = (c0 + 32 >= len(gates))
                if wide:
                    # wide stream (uint32 units): [nl | sync<<8, out[32], literal rows nl x 32]
                    chw = np.zeros(1 + 32 + 32 * nl, dtype=np.uint32)
                    chw[0] = nl | ((1 if last else 0) << 8)
                    outs32 = np.full(32, 0xFFFFFFFF, dtype=np.uint32)
                    lits32 = np.zeros((nl, 32), dtype=np.uint32)
                    for i, g in enumerate(cg):
                        outs32[i] = new_slot[g]
                        ml = [map_lit(l) for l in pn.gate_lits[g]]
                        lits32[:len(ml), i] = ml
                        n_g += 1
                        n_l += len(ml)
                    chw[1:33] = outs32
                    chw[33:] = lits32.reshape(-1)
                    chunks.append(chw)
                    continue
                nl += nl & 1
                ch = np.zeros(2 + 32 + 32 * nl, dtype=np.uint16)
                ch[0] = nl | ((1 if last else 0) << 8)
                outs = np.zeros(32, dtype=np.uint16)
                lits = np.zeros((nl, 32), dtype=np.uint16)    # literal 0 == TRUE padding
                for i in range(32):
                    if i < len(cg):
                        g = cg[i]
                        outs[i] = new_slot[g]
                        ml = [map_lit(l) for l in pn.gate_lits[g]]
                        lits[:len(ml), i] = ml
                        n_g += 1
                        n_l += len(ml)
                    else:
                        outs[i] = 0xFFFF   # patched to the dump slot below
                ch[2:34] = outs
                # pair-interleave: for pair p, lane i: [lit[2p][i], lit[2p+1][i]]
                pairs = lits.reshape(nl // 2, 2, 32).transpose(0, 2, 1).reshape(-1)
                ch[34:] = pairs
                chunks.append(ch)
            # gates whose last reader is in this level: release after the level completes.
            for g in release_at.get(li, ()):
                pending_release.append(new_slot[g])
            # gates never read and not kept (shouldn't exist) would leak a slot; harmless.
        streams.append(np.concatenate(chunks) if chunks else np.zeros(0, np.uint32 if wide else np.uint16))
        stats.append((n_g, n_l))
    kept = {}
    for l in keep_lits:
        s = int(l) >> 1
        h = out_to_gate.get(s)
        kept[s] = new_slot[h] if h is not None else s
    return streams, kept, alloc.high, stats


def _levelise_part(pn, gates, out_to_gate, prior_level, base_level):
    inpart = set(gates)
    lv: Dict[int, int] = {}
    for g in gates:
        m = base_level
        for l in pn.gate_lits[g]:
            h = out_to_gate.get(int(l) >> 1)
            if h is not None and h in inpart:
                m = max(m, lv[h])
        lv[g] = m + 1
    return lv


def _remap_lits(lits: np.ndarray, kept: Dict[int, int]) -> np.ndarray:
    out = np.zeros(len(lits), dtype=np.int64)
    for i, l in enumerate(lits):
        l = int(l)
        out[i] = kept[l >> 1] * 2 | (l & 1)
    return out


def pack_device_program(pn: Propnet, force_wide=None) -> np.ndarray:
    """Returns the int32 device blob for ``gz_propnet_create``."""
    nr = len(pn.roles)
    role_sizes = pn.role_sizes
    n_inputs = sum(role_sizes)
    nb = pn.n_bases
    sw = tables.state_words(nb)
    lwo, lw = tables.legal_layout(role_sizes)
    leaves_a = 1 + nb + n_inputs
    leaves_b = 1 + nb

    keep_b = [l for ls in pn.legal_lit for l in ls] + [pn.terminal_lit] + \
        [l for ls in pn.goal_lit for l in ls]
    max_fanin = max((len(l) for l in pn.gate_lits), default=0)
    for wide in ([False, True] if force_wide is None else [bool(force_wide)]):
        try:
            # phase A: next-state cone
            sa, kept_a, high_a, st_a = _pack_phase(pn, [pn.prog['T']], list(pn.next_lit), leaves_a, wide)
            # phase B: LT then GL in one slot space
            sb, kept_b, high_b, st_b = _pack_phase(pn, [pn.prog['LT'], pn.prog['GL']], keep_b, leaves_b, wide)
        except OverflowError:
            if wide:
                raise
            continue
        n_val = max(high_a, high_b)
        if wide or (n_val <= MAX_SLOT and max_fanin < 255):
            break
        # more live slots than a 15-bit slot index addresses: the wide format (32-bit literals, value
        # array in global memory) takes over
    dump = n_val
    val_words = n_val + 32

    def finish(stream: np.ndarray) -> Tuple[np.ndarray, int]:
        s = stream.copy()
        if wide:
            pos = 0
            nch = 0
            while pos < len(s):
                nl = int(s[pos]) & 0xFF
                outs = s[pos + 1:pos + 33]
                pad = outs == 0xFFFFFFFF
                outs[pad] = dump + np.arange(32, dtype=np.uint32)[pad]
                pos += 33 + 32 * nl
                nch += 1
            return s, nch
        # patch padding lanes' out slots to per-lane dump slots
        pos = 0
        nch = 0
        while pos < len(s):
            nl = int(s[pos]) & 0xFF
            outs = s[pos + 2:pos + 34]
            for i in range(32):
                if outs[i] == 0xFFFF:
                    outs[i] = dump + i
            pos += 34 + 32 * nl
            nch += 1
        if len(s) & 1:
            s = np.concatenate([s, np.zeros(1, np.uint16)])
        return s, nch

    pa, nca = finish(sa[0])
    plt, nclt = finish(sb[0])
    pgl, ncgl = finish(sb[1])

    next_lit = _remap_lits(pn.next_lit, kept_a)
    legal_lit = _remap_lits(np.concatenate(pn.legal_lit), kept_b)
    term_lit = int(_remap_lits(np.array([pn.terminal_lit]), kept_b)[0])
    goal_lit = _remap_lits(np.concatenate(pn.goal_lit), kept_b) if nr else np.zeros(0, np.int64)
    goal_val = np.concatenate(pn.goal_values) if nr else np.zeros(0, np.int64)
    goal_off = np.concatenate([[0], np.cumsum([len(v) for v in pn.goal_values])])

    hdr = np.zeros(DEV_HEADER_WORDS, dtype=np.int64)
    sections = []
    pos = DEV_HEADER_WORDS

    def add(arr, idx):
        nonlocal pos
        arr = np.asarray(arr).astype(np.int64).ravel()
        hdr[idx] = pos
        sections.append(arr)
        pos += len(arr)

    def add_stream(s16: np.ndarray, idx):
        add((s16 if wide else s16.view('<u4')).astype(np.int64), idx)

    add(role_sizes, D_OFF_ROLE_SIZES)
    add(lwo, D_OFF_LEGAL_WORD_OFF)
    add(tables.pack_state_bits(pn.init_state, sw).astype(np.int64), D_OFF_INIT)
    add_stream(pa, D_OFF_PROG_A)
    add_stream(plt, D_OFF_PROG_LT)
    add_stream(pgl, D_OFF_PROG_GL)
    add(next_lit, D_OFF_NEXT_LIT)
    add(legal_lit, D_OFF_LEGAL_LIT)
    add(goal_off, D_OFF_GOAL_OFF)
    add(goal_val, D_OFF_GOAL_VAL)
    add(goal_lit, D_OFF_GOAL_LIT)
    hdr[D_MAGIC] = DEV_MAGIC
    hdr[D_VERSION] = DEV_VERSION
    hdr[D_N_BASES] = nb
    hdr[D_N_ROLES] = nr
    hdr[D_N_INPUTS] = n_inputs
    hdr[D_STATE_WORDS] = sw
    hdr[D_LEGAL_WORDS] = lw
    hdr[D_TOTAL_WORDS] = pos
    hdr[D_VAL_WORDS] = val_words
    hdr[D_TERMINAL_LIT] = term_lit
    # stream lengths in 16-bit units (wide streams: 2 per word) so that words = halfs / 2 either way
    k = 2 if wide else 1
    hdr[D_CHUNKS_A], hdr[D_HALFS_A] = nca, len(pa) * k
    hdr[D_CHUNKS_LT], hdr[D_HALFS_LT] = nclt, len(plt) * k
    hdr[D_CHUNKS_GL], hdr[D_HALFS_GL] = ncgl, len(pgl) * k
    hdr[D_WIDE] = 1 if wide else 0
    hdr[D_N_GATES_A], hdr[D_N_LITS_A] = st_a[0]
    hdr[D_N_GATES_LT], hdr[D_N_LITS_LT] = st_b[0]
    hdr[D_N_GATES_GL], hdr[D_N_LITS_GL] = st_b[1]
    hdr[D_DUMP_SLOT] = dump
    blob = np.concatenate([hdr] + sections)
    return (blob & 0xFFFFFFFF).astype(np.uint32).view(np.int32)


# ------------------------------------------------------------------------------------------------
# Host-side bit-parallel emulation of the device format (used by CPU tests to validate the packer
# without a GPU; *not* a product path — the kernels in csrc/ are the product).
def emulate_stream(blob: np.ndarray, which: str, val: np.ndarray):
    """Runs program 'A' | 'LT' | 'GL' of a device blob on ``val`` (uint32 [val_words], 32 games)."""
    h = blob[:DEV_HEADER_WORDS].astype(np.int64)
    off, nch, nh = {'A': (h[D_OFF_PROG_A], h[D_CHUNKS_A], h[D_HALFS_A]),
                    'LT': (h[D_OFF_PROG_LT], h[D_CHUNKS_LT], h[D_HALFS_LT]),
                    'GL': (h[D_OFF_PROG_GL], h[D_CHUNKS_GL], h[D_HALFS_GL])}[which]
    if h[D_WIDE]:
        s = blob[off:off + nh // 2].view(np.uint32)
        pos = 0
        for _ in range(int(nch)):
            nl = int(s[pos]) & 0xFF
            outs = s[pos + 1:pos + 33].astype(np.int64)
            lits = s[pos + 33:pos + 33 + 32 * nl].reshape(nl, 32).astype(np.int64)
            acc = np.full(32, 0xFFFFFFFF, dtype=np.uint32)
            for k in range(nl):
                l = lits[k]
                acc &= val[l >> 1] ^ (np.where(l & 1, 0xFFFFFFFF, 0).astype(np.uint32))
            val[outs] = acc
            pos += 33 + 32 * nl
        return val
    s = blob[off:off + (nh + 1) // 2].view('<u2')
    pos = 0
    for _ in range(int(nch)):
        nl = int(s[pos]) & 0xFF
        outs = s[pos + 2:pos + 34].astype(np.int64)
        pairs = s[pos + 34:pos + 34 + 32 * nl].reshape(nl // 2, 32, 2).astype(np.int64)
        acc = np.full(32, 0xFFFFFFFF, dtype=np.uint32)
        for p in range(nl // 2):
            for k in range(2):
                l = pairs[p, :, k]
                v = val[l >> 1] ^ (np.where(l & 1, 0xFFFFFFFF, 0).astype(np.uint32))
                acc &= v
        val[outs] = acc
        pos += 34 + 32 * nl
    return val
