"""Board symmetries (src/ggpzero/util/symmetry.py) and the reference's own propnet property test.

CPU: src/test/test_symmetry.py:187-231 — for every symmetry T of a game and positions reached by random
play, `all_moves(sm, T(state)) == T(all_moves(sm, state))` — run against the compiled propnet through the C
oracle (this is the one asserted propnet property the reference has), plus its
known answer translate_action(role 1, action 1, no reflection, 1 rotation) == 8 for reversi (:90-91).
GPU: the batched bit-permutation kernel equals Translator.translate_basestate, and the device state machine
satisfies the same equivariance on whole batches."""
import numpy as np
import pytest

from galvanise_zero_b200 import lookup, symmetry, tables
from galvanise_zero_b200.defs import gamedesc
from oracle.pnref import RefStateMachine

GAMES = [('reversi', (5, 15)), ('breakthroughSmall', (2, 8)), ('breakthrough', (3, 12)), ('connect6', (3, 10)),
         ('amazons_10x10', (3, 10, 16)), ('hexLG13', (3, 10, 16))]


def _translator(game):
    gi = lookup.by_name(game)
    desc = getattr(gamedesc.Games(), game)()
    sym = gamedesc.get_symmetries_desc(game)
    return gi, symmetry.create_translator(gi, desc, sym), symmetry.Prescription(sym)


def _random_state(sm, plies, seed):
    s = np.zeros((1, sm.state_words), dtype=np.uint32)
    l = np.zeros((1, sm.legal_words), dtype=np.uint32)
    t = np.ones(1, dtype=np.uint8)
    g = np.zeros((1, sm.n_roles), dtype=np.uint8)
    for step in range(plies + 1):                 # first step places the initial state
        prev = s.copy()
        sm.playout_step(0, seed, step, s, l, t, g, None)
        if t[0]:
            return prev[0]
    return s[0]


def _all_moves(sm, state_words):
    sm.update_bases(state_words)
    return [sorted(sm.legals(r)) for r in range(sm.n_roles)]


def test_reversi_known_answer_and_prescriptions():
    gi, t, pres = _translator('reversi')
    assert t.translate_action(1, 1, False, 1) == 8          # src/test/test_symmetry.py:90-91
    assert list(pres) == [(False, r) for r in range(4)] + [(True, r) for r in range(4)]
    assert list(_translator('hexLG13')[2]) == [(False, 0), (False, 2)]
    assert list(_translator('breakthrough')[2]) == [(False, 0), (True, 0)]


@pytest.mark.parametrize('game,advances', GAMES)
def test_legal_moves_are_equivariant_under_every_symmetry(game, advances):
    gi, t, pres = _translator(game)
    sm = RefStateMachine(tables.blob_by_name(game))
    nb = len(gi.model.bases)
    for plies in advances:
        state = _random_state(sm, plies, seed=100 + plies)
        bits = [int(x) for x in tables.unpack_state_bits(state, nb)]
        moves = _all_moves(sm, state)
        assert any(len(m) > 1 for m in moves)
        for do_reflection, rot_count in pres:
            tbits = t.translate_basestate(bits, do_reflection, rot_count)
            assert sum(tbits) == sum(bits)
            tstate = tables.pack_state_bits(np.array(tbits, dtype=np.uint8), sm.state_words)
            want = [sorted(t.translate_action(r, a, do_reflection, rot_count) for a in moves[r])
                    for r in range(sm.n_roles)]
            assert _all_moves(sm, tstate) == want, (game, plies, do_reflection, rot_count)
            # the permutation tables used by the device path say the same
            dest = t.base_permutation(do_reflection, rot_count)
            via = np.zeros(nb, dtype=np.uint8)
            via[dest[np.nonzero(bits)[0]]] = 1
            assert via.tolist() == tbits


@pytest.mark.gpu
@pytest.mark.parametrize('game', ['reversi', 'amazons_10x10', 'hexLG13'])
def test_device_symmetry_of_state_batches(game):
    import torch
    from galvanise_zero_b200.statemachine import DeviceStateMachine
    gi, t, pres = _translator(game)
    ref = RefStateMachine(tables.blob_by_name(game))
    dev = DeviceStateMachine(game)
    n = 70 if game != 'hexLG13' else 12
    nb = len(gi.model.bases)
    s = np.zeros((n, ref.state_words), dtype=np.uint32)
    l = np.zeros((n, ref.legal_words), dtype=np.uint32)
    tm = np.ones(n, dtype=np.uint8)
    g = np.zeros((n, ref.n_roles), dtype=np.uint8)
    for step in range(9):
        ref.playout_step(0, 77, step, s, l, tm, g, None)
    sd = torch.from_numpy(s.view(np.int32)).to(dev.device)
    legal0 = dev.eval(sd)[0].cpu().numpy().view(np.uint32)
    for do_reflection, rot_count in pres:
        ds = symmetry.DeviceSymmetry(t, do_reflection, rot_count)
        out = ds.states(sd)
        got = out.cpu().numpy().view(np.uint32)
        for i in range(n):
            bits = [int(x) for x in tables.unpack_state_bits(s[i], nb)]
            want = tables.pack_state_bits(np.array(t.translate_basestate(bits, do_reflection, rot_count), dtype=np.uint8),
                                          ref.state_words)
            assert np.array_equal(got[i], want)
        # equivariance on the device: legal masks of the transformed batch = transformed legal masks
        legal_t = dev.eval(out)[0].cpu().numpy().view(np.uint32)
        for r in range(ref.n_roles):
            a_n = dev.role_sizes[r]
            w0, w1 = dev.legal_word_off[r], dev.legal_word_off[r + 1]
            m0 = tables.unpack_state_bits(legal0[:, w0:w1], a_n)
            mt = tables.unpack_state_bits(legal_t[:, w0:w1], a_n)
            dest = t.action_permutation(r, do_reflection, rot_count)
            moved = np.zeros_like(m0)
            moved[:, dest] = m0
            assert np.array_equal(mt, moved)
            dense = torch.from_numpy(m0.astype(np.float32)).to(dev.device)
            assert np.array_equal(ds.policy(r, dense).cpu().numpy(), moved.astype(np.float32))
