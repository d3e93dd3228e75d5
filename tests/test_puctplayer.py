"""PUCTPlayer (src/ggpzero/player/puctplayer.py:13-108): the caller of the Player path.

CPU: the hooks' logic against stand-ins for the state machine and the poller — who is on move is read from the
noop legals, `playouts_per_iteration` when it is our move and `playouts_per_iteration_noop` otherwise, the move /
probability / node count come from `player_get_move(our_role_index)`, every joint move is applied and polled.
GPU: two PUCTPlayers play breakthroughSmall to the end through the real PlayPoller / Player / device engine with
the synthetic network (every move legal on the referee's state machine, goals as the rules say, the same game when
replayed), and the searches a player runs are the ones the parity tests of tests/test_gpu_api.py pin to the
reference's evaluator (same PlayPoller calls)."""
import pytest

from galvanise_zero_b200 import puctplayer
from galvanise_zero_b200.defs import confs


class _Legals(object):
    def __init__(self, xs):
        self.xs = xs

    def get_count(self):
        return len(self.xs)

    def get_legal(self, i):
        return self.xs[i]


class _FakeSM(object):
    def __init__(self, legals_by_state):
        self.legals_by_state = legals_by_state
        self.updated = []

    def update_bases(self, state):
        self.cur = state
        self.updated.append(state)

    def get_legal_state(self, ri):
        return _Legals(self.legals_by_state[self.cur][ri])


class _FakePoller(object):
    def __init__(self):
        self.calls = []

    def player_reset(self, depth):
        self.calls.append(('reset', depth))

    def player_move(self, state, iterations, finish_time):
        self.calls.append(('move', state, iterations, finish_time))

    def player_apply_move(self, jm):
        self.calls.append(('apply', jm))

    def poll_loop(self):
        self.calls.append(('poll_loop',))

    def player_get_move(self, role):
        self.calls.append(('get_move', role))
        return 17 + role, 0.625, 4242

    def player_balance_moves(self, n):
        self.calls.append(('balance', n))

    def player_tree_debug(self, n):
        return (('dbg', n),)

    def player_update_config(self, *a, **k):
        self.calls.append(('update_config', a, k))


class _FakeMatch(object):
    match_id = 'm1'
    game_depth = 3

    def __init__(self, role, state):
        self.our_role_index = role
        self.state = state

    def get_current_state(self):
        return self.state


def _player(role, state):
    conf = confs.PUCTPlayerConfig(name='p', playouts_per_iteration=321, playouts_per_iteration_noop=5, generation='g7')
    p = puctplayer.PUCTPlayer(conf)
    assert p.get_name() == 'p_321_g7' and repr(p) == 'p_321_g7'
    assert puctplayer.PUCTPlayer(confs.PUCTPlayerConfig(name='q', playouts_per_iteration=-1, generation='x')).get_name() == 'q_x'
    # role 0's noop is legal 0, role 1's noop is legal 9; in state 'A' role 0 moves, in state 'B' role 1 moves
    p.sm = _FakeSM({'A': [[3, 4], [9]], 'B': [[0], [2, 5, 6]]})
    p.role0_noop_legal, p.role1_noop_legal = 0, 9
    p.poller = _FakePoller()
    p.reset(_FakeMatch(role, state))
    return p


def test_hooks_follow_the_reference_call_sequence():
    p = _player(0, 'A')                                        # our move
    assert p.on_next_move(12.5) == 17
    assert p.poller.calls == [('move', 'A', 321, 12.5), ('poll_loop',), ('get_move', 0)]
    assert p.last_probability == 0.625 and p.last_node_count == 4242 and p.sm.updated == ['A']
    q = _player(1, 'A')                                        # the opponent's move: the noop budget, our role's score
    assert q.on_next_move(-1) == 18
    assert q.poller.calls == [('move', 'A', 5, -1), ('poll_loop',), ('get_move', 1)]
    r = _player(1, 'B')
    assert r.lead_role_index() == 1 and r.on_next_move(-1) == 18
    assert r.poller.calls[0] == ('move', 'B', 321, -1)
    r.poller.calls.clear()
    r.on_apply_move('jm')
    r.balance_moves(3)
    r.update_config(2.0, converged_visits=100)
    r.cleanup()
    assert r.poller.calls == [('apply', 'jm'), ('poll_loop',), ('balance', 3), ('poll_loop',),
                              ('update_config', (2.0,), {'converged_visits': 100}), ('reset', 0)]
    assert r.tree_debug(4) == (('dbg', 4),)
    bad = _player(0, 'A')
    bad.sm = _FakeSM({'A': [[3, 4], [2, 5]]})                  # nobody has a lone noop: not a turn-taking position
    with pytest.raises(AssertionError):
        bad.on_next_move(-1)


def test_manager_registry():
    m = puctplayer.get_manager()
    assert m is puctplayer.get_manager()
    sentinel = object()
    m.register('reversi', 'gen_x', sentinel)
    assert m.load_network('reversi', 'gen_x') is sentinel


@pytest.mark.gpu
def test_two_puct_players_play_a_game_to_the_end():
    from galvanise_zero_b200 import bases, lookup
    from galvanise_zero_b200.defs import datadesc
    from test_gpu_api import _SynthNN
    game = 'breakthroughSmall'
    gi = lookup.by_name(game)
    tr = bases.GdlBasesTransformer(gi, datadesc.generation_for(game))
    puctplayer.get_manager().register(game, 'synth', _SynthNN(tr))

    def players():
        ev = confs.base_puct_config(dirichlet_noise_pct=-1, choose='choose_top_visits', batch_size=8)
        return [puctplayer.PUCTPlayer(confs.PUCTPlayerConfig(name='w', generation='synth', playouts_per_iteration=120,
                                                            playouts_per_iteration_noop=8, evaluator_config=ev),
                                      seed=0x5EED0000 + k) for k in range(2)]

    ps = players()
    moves, goals = puctplayer.play_game(gi, ps)
    assert goals is not None and sorted(goals) == [0, 100] and 10 <= len(moves) <= 120
    noop = [next(i for i, a in enumerate(acts) if 'noop' in a) for acts in gi.model.actions]
    for k, jm in enumerate(moves):                             # the players alternate, the other role passes
        assert jm[1 - k % 2] == noop[1 - k % 2] and jm[k % 2] != noop[k % 2]
    assert all(p.last_node_count > 0 and -0.06 <= p.last_probability <= 1.06 for p in ps)
    moves2, goals2 = puctplayer.play_game(gi, players())
    assert moves2 == moves and goals2 == goals                 # same seeds, same network: the same game
    cut, g = puctplayer.play_game(gi, players(), max_moves=4)
    assert cut == moves[:4] and g is None
