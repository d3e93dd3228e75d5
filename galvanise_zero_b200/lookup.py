"""Game lookup: ``by_name(game) -> GameInfo`` with ``.game``, ``.model`` and ``.get_sm()``.

Mirrors the part of ``ggplib.db.lookup`` the reference uses (``src/ggpzero/nn/bases.py:26-29,79-84``,
``src/ggpzero/util/symmetry.py:47,58-59``, ``src/test/cpp/test_interface.py:40-44``): bases are
strings like ``(true (control black))`` and actions ``(does black (move 1 1))``.
"""
from functools import lru_cache

from . import tables


class GameModel(object):
    def __init__(self, pn):
        self.roles = list(pn.roles)
        self.bases = ['(true %s)' % b for b in pn.bases]
        self.actions = [['(does %s %s)' % (r, m) for m in pn.inputs[ri]]
                        for ri, r in enumerate(pn.roles)]

    def basestate_to_str(self, bs):
        return ' '.join(self.bases[i] for i in range(len(self.bases)) if bs.get(i))


class GameInfo(object):
    def __init__(self, game):
        self.game = game
        self.propnet = tables.propnet_by_name(game)
        self.model = GameModel(self.propnet)

    def get_sm(self, device=0):
        from .ggplib_api import StateMachine
        return StateMachine(self, device=device)


@lru_cache(maxsize=None)
def by_name(game):
    return GameInfo(game)
