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


def test_fixtures_cover_the_rules_corner_cases():
    """The golden trajectories (>= 20 games per game played to their end; hexLG13: 6) contain the situations a
    propnet is most likely to get wrong, so test_oracle_matches_golden / the GPU propnet tests pin them:
    reversi forced passes and both kinds of game end, the hex swap, chess castling on both wings, en passant,
    promotions, check evasions and the step counter running out."""
    fx = load_fixture('reversi')
    assert len(fx['games']) >= 20
    passes = ended = 0
    for g in fx['games']:
        ended += int(g[-1]['terminal'])
        for rec in g:
            if not rec['terminal'] and all(len(l) == 1 for l in rec['legal']):
                passes += 1           # the side to move has no stone to place: both roles can only noop
    assert ended >= 20 and passes >= 3

    fx = load_fixture('hexLG13')
    swap = [i for i, m in enumerate(fx['inputs'][fx['roles'].index('white')]) if m == 'swap']
    assert len(swap) == 1
    wi = fx['roles'].index('white')
    swaps = sum(1 for g in fx['games'] for rec in g if rec['move'] is not None and rec['move'][wi] == swap[0])
    offered = sum(1 for g in fx['games'] for rec in g if swap[0] in rec['legal'][wi])
    assert swaps >= 2 and offered >= len(fx['games']) and sum(int(g[-1]['terminal']) for g in fx['games']) >= 4

    fx = load_fixture('chess_15d')
    assert len(fx['games']) >= 20
    bases, inputs = fx['bases'], fx['inputs']
    ev = dict(castle_k=0, castle_q=0, en_passant=0, promotion=0, step32=0, check=0, terminal=0)
    for g in fx['games']:
        ev['terminal'] += int(g[-1]['terminal'])
        for rec in g:
            st = set(bases[i] for i in rec['state'])
            ev['step32'] += int('(step 32)' in st)
            if rec['move'] is None:
                continue
            for r in range(2):
                m = inputs[r][rec['move'][r]]
                ev['castle_k'] += int(m == 'castleKingside')
                ev['castle_q'] += int(m == 'castleQueenside')
                t = m.strip('()').split()
                if t[0] != 'move':
                    continue
                if t[1] == 'pawn' and t[2] != t[4] and not any(b.startswith('(cell %s %s ' % (t[4], t[5])) for b in st):
                    ev['en_passant'] += 1          # a pawn captures onto an empty square
                if t[1] != 'pawn' and ('(cell %s %s white pawn)' % (t[2], t[3]) in st or
                                       '(cell %s %s black pawn)' % (t[2], t[3]) in st):
                    ev['promotion'] += 1           # the piece that moves is a pawn and arrives as something else
                if len(rec['legal'][r]) <= 3 and len(rec['legal'][r]) > 1 and '(step 1)' not in st:
                    ev['check'] += 1               # almost no legal move: check evasion (or a bare king)
    assert ev['castle_k'] >= 2 and ev['castle_q'] >= 2 and ev['en_passant'] >= 2, ev
    assert ev['promotion'] >= 5 and ev['step32'] >= 1 and ev['check'] >= 5 and ev['terminal'] >= 3, ev

    for game in ('breakthrough', 'breakthroughSmall', 'connect6', 'amazons_10x10'):
        fx = load_fixture(game)
        assert len(fx['games']) >= 20, game
    for game in ('amazons_10x10', 'connect6'):
        assert sum(int(g[-1]['terminal']) for g in load_fixture(game)['games']) >= 20, game


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
