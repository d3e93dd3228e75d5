"""``PUCTPlayer`` — the caller of the Player path, mirroring ``src/ggpzero/player/puctplayer.py:13-108``: same
class, attributes, hooks (``on_meta_gaming`` / ``on_apply_move`` / ``on_next_move`` / ``cleanup``) and the same
call sequence into ``PlayPoller`` (``player_reset`` -> per move ``player_move`` + ``poll_loop`` +
``player_get_move``; ``player_apply_move`` + ``poll_loop`` after every joint move).

What the reference takes from ggplib and this build has no copy of (ggplib is an external dependency, SURVEY.md
§8(c)) is reduced to the surface ``PUCTPlayer`` touches: ``MatchPlayer`` (``ggplib.player.base``: name, ``match``,
the hooks) and ``Match`` (``ggplib.player.match``: ``match_id``, ``game_info``, ``game_depth``, ``our_role_index``,
``get_current_state()``), plus a small ``play_game`` driver standing where ggplib's ``GameMaster`` runs players
against each other (``src/test/player/test_player.py:30-75``). The network comes from a manager with the reference's
``get_manager().load_network(game, generation)`` call (``src/ggpzero/nn/manager.py``); there being no checkpoint format
here, networks are registered on it, or built with random weights for a generation nobody registered."""
from .cppinterface import PlayPoller
from .defs import confs


def joint_move_to_ptr(joint_move):
    """``cppinterface.joint_move_to_ptr`` (``src/ggpzero/util/cppinterface.py:25-26``): the engine objects of this
    build take the Python object or its address alike."""
    return joint_move


def basestate_to_ptr(basestate):
    """``cppinterface.basestate_to_ptr`` (``src/ggpzero/util/cppinterface.py:29-30``)."""
    return basestate


class NetworkManager(object):
    """``get_manager().load_network(game, generation)`` (``src/ggpzero/nn/manager.py``) over a registry."""

    def __init__(self):
        self._nets = {}

    def register(self, game, generation, nn):
        self._nets[(game, generation)] = nn

    def load_network(self, game, generation):
        key = (game, generation)
        if key not in self._nets:
            from . import bases, lookup
            from . import nn as gznn
            from .defs import datadesc
            tr = bases.GdlBasesTransformer(lookup.by_name(game), datadesc.generation_for(game))
            self._nets[key] = gznn.NeuralNetwork(tr, gznn.make_inference_network(tr), datadesc.generation_for(game))
        return self._nets[key]


_manager = NetworkManager()


def get_manager():
    return _manager


class MatchPlayer(object):
    """The part of ``ggplib.player.base.MatchPlayer`` a player implements / a game master calls."""

    def __init__(self, name=None):
        self.name = name or self.__class__.__name__
        self.match = None

    def get_name(self):
        return self.name

    def reset(self, match):
        self.match = match

    def cleanup(self):
        pass

    def on_meta_gaming(self, finish_time):
        pass

    def on_apply_move(self, joint_move):
        pass

    def on_next_move(self, finish_time):
        raise NotImplementedError


class Match(object):
    """The part of ``ggplib.player.match.Match`` ``PUCTPlayer`` reads. One per player (``our_role_index``)."""

    def __init__(self, game_info, match_id, role_index, sm=None):
        self.game_info = game_info
        self.match_id = match_id
        self.our_role_index = role_index
        self.sm = sm or game_info.get_sm()
        self.states = [self.sm.get_initial_state()]
        self.joint_moves = []

    @property
    def game_depth(self):
        return len(self.states) - 1

    def get_current_state(self):
        return self.states[-1]

    def apply_move(self, joint_move, next_state):
        self.joint_moves.append(joint_move)
        self.states.append(next_state)


class PUCTPlayer(MatchPlayer):
    """Match player over ``PlayPoller``. Public surface as in the reference (``poller``, ``last_probability``,
    ``last_node_count``, ``identifier``, the hooks, ``balance_moves`` / ``tree_debug`` / ``update_config``); the body
    is organised around two small tables — every role's ``noop`` legal and the playout budget by "is it our move"."""
    poller = None
    last_probability = -1
    last_node_count = -1

    def __init__(self, conf, **engine_kwds):
        if not isinstance(conf, (confs.PUCTPlayerConfig, confs.PUCTEvaluatorConfig)):
            raise TypeError('PUCTPlayer wants a PUCTPlayerConfig')
        self.conf = conf
        parts = [conf.name] + ([conf.playouts_per_iteration] if conf.playouts_per_iteration > 0 else []) + [conf.generation]
        self.identifier = '_'.join(str(x) for x in parts)          # name[_playouts]_generation (puctplayer.py:21-24)
        super().__init__(self.identifier)
        self.sm = self.nn = None
        self.noop_legals = None                   # per role: the legal index of its noop
        self._engine_kwds = engine_kwds           # (seed, node_cap ... of ggpzero_interface.Player; none in the reference)

    # ---- set-up: a state machine, the generation's network and the poller, once per game type ("*" in the
    # generation name: the newest one, looked up again for every match — puctplayer.py:35-60)
    def _build(self):
        game_info = self.match.game_info
        self.sm = game_info.get_sm()
        self.nn = get_manager().load_network(game_info.game, self.conf.generation)
        self.poller = PlayPoller(self.sm, self.nn, self.conf.evaluator_config, **self._engine_kwds)
        self.noop_legals = []
        for actions in game_info.model.actions:
            found = [i for i, a in enumerate(actions) if 'noop' in a]
            if not found:
                raise ValueError('did not find noop among a role\'s actions')
            self.noop_legals.append(found[0])
        # (the reference keeps them as two attributes)
        self.role0_noop_legal, self.role1_noop_legal = self.noop_legals[0], self.noop_legals[1]

    def on_meta_gaming(self, finish_time):
        if self.sm is None or '*' in self.conf.generation:
            self._build()
        self.poller.player_reset(self.match.game_depth)

    def cleanup(self):
        if self.poller is not None:
            self.poller.player_reset(0)

    def on_apply_move(self, joint_move):
        self.poller.player_apply_move(joint_move_to_ptr(joint_move))
        self.poller.poll_loop()

    # ---- who is on move: the first role (of two) that is NOT reduced to its lone noop; the other one must be
    # (puctplayer.py:68-80 asserts the same of a turn-taking game)
    def _only_noop(self, ri):
        legals = self.sm.get_legal_state(ri)
        noops = self.noop_legals if self.noop_legals is not None else (self.role0_noop_legal, self.role1_noop_legal)
        return legals.get_count() == 1 and legals.get_legal(0) == noops[ri]

    def lead_role_index(self):
        self.sm.update_bases(self.match.get_current_state())
        if self._only_noop(0):
            return 1
        if not self._only_noop(1):
            raise AssertionError('not a turn-taking position: neither role is reduced to noop')
        return 0

    def on_next_move(self, finish_time):
        ours = self.match.our_role_index
        budget = {True: self.conf.playouts_per_iteration, False: self.conf.playouts_per_iteration_noop}
        iterations = budget[self.lead_role_index() == ours]
        self.poller.player_move(basestate_to_ptr(self.match.get_current_state()), iterations, finish_time)
        self.poller.poll_loop()
        move, self.last_probability, self.last_node_count = self.poller.player_get_move(ours)
        return move

    def balance_moves(self, max_count):
        self.poller.player_balance_moves(max_count)
        self.poller.poll_loop()

    def tree_debug(self, max_count):
        return self.poller.player_tree_debug(max_count)

    def update_config(self, *args, **kwds):
        self.poller.player_update_config(*args, **kwds)

    def __repr__(self):
        return self.get_name()


def play_game(game_info, players, match_id="b200_match", max_moves=None, move_time=-1.0):
    """Runs one game between ``players`` (one per role, in role order), the way ggplib's game master drives
    ``MatchPlayer`` objects: ``reset(match)`` + ``on_meta_gaming``, then per ply every player's ``on_next_move``,
    the joint move applied on the referee's state machine and reported through ``on_apply_move``.
    Returns ``(joint moves as legal indices, goal values per role or None when cut short)``."""
    import time
    sm = game_info.get_sm()
    matches = [Match(game_info, match_id, ri, sm=sm.dupe()) for ri in range(len(players))]
    for p, m in zip(players, matches):
        p.reset(m)
        p.on_meta_gaming(-1.0)
    state = sm.get_initial_state()
    played = []
    while True:
        sm.update_bases(state)
        if sm.is_terminal():
            goals = [sm.get_goal_value(ri) for ri in range(len(players))]
            break
        if max_moves is not None and len(played) >= max_moves:
            goals = None
            break
        finish = time.time() + move_time if move_time > 0 else -1.0
        jm = sm.get_joint_move()
        for ri, p in enumerate(players):
            mv = p.on_next_move(finish)
            legals = sm.get_legal_state(ri).to_list()
            assert mv in legals, "player %s chose an illegal move %s (legal: %s)" % (p, mv, legals)
            jm.set(ri, mv)
        nxt = sm.new_base_state()
        sm.next_state(jm, nxt)
        played.append([jm.get(ri) for ri in range(len(players))])
        for p, m in zip(players, matches):
            m.apply_move(jm, nxt)
            p.on_apply_move(jm)
        state = nxt
    for p in players:
        p.cleanup()
    return played, goals
