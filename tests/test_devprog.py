"""CPU: the device program packer (slot re-use, chunking, literal interleave) is validated against
the C oracle by emulating the device stream bit-parallel in numpy (32 games per word)."""
import numpy as np
import pytest

from galvanise_zero_b200 import devprog as dp
from galvanise_zero_b200 import tables
from oracle.pnref import RefStateMachine

GAMES = ['breakthrough', 'reversi', 'connect6', 'amazons_10x10']


def slices_from_states(states, n_bits):
    """states uint32 [32, words] -> slice words [n_bits] (bit g of word b = bit b of game g)."""
    bits = tables.unpack_state_bits(states, n_bits).astype(np.uint64)   # [32, n_bits]
    w = (bits << np.arange(32, dtype=np.uint64)[:, None]).sum(axis=0)
    return w.astype(np.uint32)


def lit_words(val, lits):
    lits = np.asarray(lits, dtype=np.int64)
    return val[lits >> 1] ^ np.where(lits & 1, 0xFFFFFFFF, 0).astype(np.uint32)


@pytest.mark.parametrize('game', GAMES + ['breakthrough:wide', 'chess_15d'])
def test_device_stream_emulation_matches_oracle(game):
    force_wide = None
    if game.endswith(':wide'):              # the wide (32-bit literal) stream format on a small game
        game, force_wide = game.split(':')[0], True
    pn = tables.propnet_by_name(game)
    blob = dp.pack_device_program(pn, force_wide=force_wide)
    if game == 'chess_15d' or force_wide:
        assert blob[dp.D_WIDE] == 1
    h = blob[:dp.DEV_HEADER_WORDS].astype(np.int64)
    sm = RefStateMachine(tables.blob_by_name(game))
    nb, nr = pn.n_bases, len(pn.roles)
    sizes = pn.role_sizes
    n = 32
    next_lit = blob[h[dp.D_OFF_NEXT_LIT]:h[dp.D_OFF_NEXT_LIT] + nb]
    legal_lit = blob[h[dp.D_OFF_LEGAL_LIT]:h[dp.D_OFF_LEGAL_LIT] + sum(sizes)]
    goal_off = blob[h[dp.D_OFF_GOAL_OFF]:h[dp.D_OFF_GOAL_OFF] + nr + 1]
    goal_val = blob[h[dp.D_OFF_GOAL_VAL]:h[dp.D_OFF_GOAL_VAL] + goal_off[-1]]
    goal_lit = blob[h[dp.D_OFF_GOAL_LIT]:h[dp.D_OFF_GOAL_LIT] + goal_off[-1]]
    lwo = blob[h[dp.D_OFF_LEGAL_WORD_OFF]:h[dp.D_OFF_LEGAL_WORD_OFF] + nr + 1]

    states = np.zeros((n, sm.state_words), dtype=np.uint32)
    legal = np.zeros((n, sm.legal_words), dtype=np.uint32)
    term = np.ones(n, dtype=np.uint8)
    goals = np.zeros((n, nr), dtype=np.uint8)
    moves = np.zeros((n, nr), dtype=np.int16)
    val = np.zeros(int(h[dp.D_VAL_WORDS]), dtype=np.uint32)
    steps = {'amazons_10x10': 40, 'chess_15d': 6}.get(game, 70)
    for step in range(steps):
        prev = states.copy()
        was_term = term.copy()
        sm.playout_step(0, 7, step, states, legal, term, goals, moves)
        live = was_term == 0
        if not live.any():
            continue
        # phase A on prev states with the oracle's chosen moves
        val[:] = 0
        val[0] = 0xFFFFFFFF
        val[1:1 + nb] = slices_from_states(prev, nb)
        off = 1 + nb
        for r in range(nr):
            for g in range(n):
                if live[g]:
                    val[off + int(moves[g, r])] |= np.uint32(1 << g)
            off += sizes[r]
        dp.emulate_stream(blob, 'A', val)
        nxt_slices = lit_words(val, next_lit)
        want = slices_from_states(states, nb)
        mask = np.uint32(sum(1 << g for g in range(n) if live[g]))
        assert np.array_equal(nxt_slices & mask, want & mask)
        # phase B on the new states
        val[:] = 0
        val[0] = 0xFFFFFFFF
        val[1:1 + nb] = want
        dp.emulate_stream(blob, 'LT', val)
        dp.emulate_stream(blob, 'GL', val)
        tw = lit_words(val, [int(h[dp.D_TERMINAL_LIT])])[0]
        lw = lit_words(val, legal_lit)
        for g in range(n):
            assert ((int(tw) >> g) & 1) == term[g]
            o = 0
            for r in range(nr):
                bits = (lw[o:o + sizes[r]] >> np.uint32(g)) & 1
                got = tables.pack_state_bits(bits.astype(np.uint8), (sizes[r] + 31) // 32)
                assert np.array_equal(got, legal[g, lwo[r]:lwo[r + 1]])
                o += sizes[r]
                gv = 255
                for k in range(goal_off[r], goal_off[r + 1]):
                    if (int(lit_words(val, [goal_lit[k]])[0]) >> g) & 1:
                        gv = int(goal_val[k])
                assert gv == goals[g, r]


def test_slot_reuse_keeps_value_array_small():
    pn = tables.propnet_by_name('reversi')
    blob = dp.pack_device_program(pn)
    assert blob[dp.D_VAL_WORDS] * 4 < pn.n_slots
