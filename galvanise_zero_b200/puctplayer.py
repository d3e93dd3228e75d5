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
    poller = None
    last_probability = -1
    last_node_count = -1

    def __init__(self, conf, **engine_kwds):
        assert isinstance(conf, (confs.PUCTPlayerConfig, confs.PUCTEvaluatorConfig))
        self.conf = conf
        if conf.playouts_per_iteration > 0:
            self.identifier = "%s_%s_%s" % (self.conf.name, conf.playouts_per_iteration, conf.generation)
        else:
            self.identifier = "%s_%s" % (self.conf.name, conf.generation)
        super().__init__(self.identifier)
        self.sm = None
        self._engine_kwds = engine_kwds          # (seed, node_cap ... of ggpzero_interface.Player; none in the reference)

    def cleanup(self):
        if self.poller is not None:
            self.poller.player_reset(0)

    def on_meta_gaming(self, finish_time):
        if self.sm is None or "*" in self.conf.generation:
            game_info = self.match.game_info
            self.sm = game_info.get_sm()
            self.nn = get_manager().load_network(game_info.game, self.conf.generation)
            self.poller = PlayPoller(self.sm, self.nn, self.conf.evaluator_config, **self._engine_kwds)

            def get_noop_idx(actions):
                for idx, a in enumerate(actions):
                    if "noop" in a:
                        return idx
                assert False, "did not find noop"

            self.role0_noop_legal, self.role1_noop_legal = map(get_noop_idx, game_info.model.actions)
        self.poller.player_reset(self.match.game_depth)

    def on_apply_move(self, joint_move):
        self.poller.player_apply_move(joint_move_to_ptr(joint_move))
        self.poller.poll_loop()

    def lead_role_index(self):
        """Who is on move in the current state: the role whose only legal is not ``noop`` (puctplayer.py:68-80)."""
        current_state = self.match.get_current_state()
        self.sm.update_bases(current_state)
        if (self.sm.get_legal_state(0).get_count() == 1 and
                self.sm.get_legal_state(0).get_legal(0) == self.role0_noop_legal):
            return 1
        assert (self.sm.get_legal_state(1).get_count() == 1 and
                self.sm.get_legal_state(1).get_legal(0) == self.role1_noop_legal)
        return 0

    def on_next_move(self, finish_time):
        lead_role_index = self.lead_role_index()
        if lead_role_index == self.match.our_role_index:
            max_iterations = self.conf.playouts_per_iteration
        else:
            max_iterations = self.conf.playouts_per_iteration_noop
        current_state = self.match.get_current_state()
        self.poller.player_move(basestate_to_ptr(current_state), max_iterations, finish_time)
        self.poller.poll_loop()
        move, prob, node_count = self.poller.player_get_move(self.match.our_role_index)
        self.last_probability = prob
        self.last_node_count = node_count
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
