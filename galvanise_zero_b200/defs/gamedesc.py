"""Per-game plane descriptions: which bases become board channels and control channels.

Same descriptor vocabulary as ``src/ggpzero/defs/gamedesc.py:5-70`` (ControlBase, ControlChannel,
BoardTerm, BoardChannels, GameDesc) so that descriptions written for the reference load unchanged;
the games listed are the ones named by BASELINE.json's configs plus the small test games
(values from ``gamedesc.py:142-160,173-180,216-232,309-318,320-364,430-440``).
"""
import attr


@attr.s
class ControlBase(object):
    arg_terms = attr.ib(default=attr.Factory(list))   # must match the base's terms exactly
    value = attr.ib(default=1)                        # the whole channel is set to this


@attr.s
class ControlChannel(object):
    control_bases = attr.ib(default=attr.Factory(list))   # mutually exclusive bases


@attr.s
class BoardTerm(object):
    term_idx = attr.ib(default=3)
    terms = attr.ib(default=attr.Factory(list))


@attr.s
class BoardChannels(object):
    base_term = attr.ib(default="cell")
    x_term_idx = attr.ib(default=1)
    y_term_idx = attr.ib(default=2)
    board_terms = attr.ib(default=attr.Factory(list))     # cross product -> channels


@attr.s
class GameDesc(object):
    game = attr.ib(default="checkers")
    x_cords = attr.ib(default=attr.Factory(list))
    y_cords = attr.ib(default=attr.Factory(list))
    board_channels = attr.ib(default=attr.Factory(list))
    control_channels = attr.ib(default=attr.Factory(list))


def simple_control(*terms):
    return ControlChannel([ControlBase(list(terms), 1)])


def binary_control(base_term, a_term, b_term):
    return ControlChannel([ControlBase([base_term, a_term], 0), ControlBase([base_term, b_term], 1)])


def step_control(base_term, start, end):
    divisor = float(end - start + 1)
    return ControlChannel([ControlBase([base_term, str(i)], i / divisor)
                           for i in range(start, end + 1)])


def simple_board_channels(base, pieces):
    return BoardChannels(base, 1, 2, [BoardTerm(3, list(pieces))])


def _nums(n):
    return [str(i) for i in range(1, n + 1)]


def _letters(n):
    return list('abcdefghijklmnopqrstuvwxyz'[:n])


class Games(object):
    """Namespace: one method per game name returning its GameDesc."""

    def breakthrough(self):
        return GameDesc("breakthrough", _nums(8), _nums(8),
                        [simple_board_channels("cellHolds", ["white", "black"])],
                        [binary_control("control", "black", "white")])

    def breakthroughSmall(self):
        return GameDesc("breakthroughSmall", _nums(6), _nums(6),
                        [simple_board_channels("cell", ["white", "black"])],
                        [binary_control("control", "white", "black")])

    def reversi(self):
        return GameDesc("reversi", _nums(8), _nums(8),
                        [simple_board_channels("cell", ["black", "red"])],
                        [binary_control("control", "black", "red")])

    def connectFour(self):
        return GameDesc("connectFour", _nums(8), ["6", "5", "4", "3", "2", "1"],
                        [simple_board_channels("cell", ["red", "black"])],
                        [binary_control("control", "red", "black")])

    def hexLG11(self):
        return GameDesc("hex", _letters(11), _nums(11),
                        [simple_board_channels("cell", ["black", "white"])],
                        [binary_control("control", "black", "white")])

    def hexLG13(self):
        return GameDesc("hex", _letters(13), _nums(13),
                        [simple_board_channels("cell", ["black", "white"])],
                        [binary_control("control", "black", "white")])

    def amazons_10x10(self):
        controls = [simple_control("turn", "black", "move"), simple_control("turn", "black", "fire"),
                    simple_control("turn", "white", "move"), simple_control("turn", "white", "fire")]
        return GameDesc("amazons_10x10", _nums(10), _nums(10),
                        [BoardChannels("justMoved", 1, 2),
                         simple_board_channels("cell", ["white", "black", "arrow"])], controls)

    def connect6(self):
        controls = [simple_control("control", t) for t in
                    ("black_turn0", "black_turn1", "white_turn0", "white_turn1")]
        return GameDesc("connect6", _nums(19), _nums(19),
                        [simple_board_channels("cell", ["black", "white"])], controls)

    def chess_15d(self):
        controls = [binary_control("control", "white", "black")]
        for s in ("kingHasMoved", "hRookHasMoved", "aRookHasMoved"):
            controls.append(simple_control(s, "black"))
            controls.append(simple_control(s, "white"))
        controls.append(step_control("step", 1, 32))
        cell = BoardChannels("cell", 1, 2, [BoardTerm(3, ["white", "black"]),
                                            BoardTerm(4, ["king", "rook", "pawn", "knight",
                                                          "queen", "bishop"])])
        return GameDesc("chess_15d", _letters(8), _nums(8), [cell], controls)


# ---- symmetry descriptions (src/ggpzero/defs/gamedesc.py:78-110,497-593)
@attr.s
class ApplySymmetry(object):
    base_term = attr.ib(default="cell")
    x_terms_idx = attr.ib(default=attr.Factory(list))     # index/indices of the terms holding x / y coordinates
    y_terms_idx = attr.ib(default=attr.Factory(list))


@attr.s
class Symmetries(object):
    apply_bases = attr.ib(default=attr.Factory(list))
    skip_bases = attr.ib(default=attr.Factory(list))
    apply_actions = attr.ib(default=attr.Factory(list))
    skip_actions = attr.ib(default=attr.Factory(list))
    do_reflection = attr.ib(default=False)           # (field order as in the reference: positional construction)
    do_rotations_90 = attr.ib(default=False)
    do_rotations_180 = attr.ib(default=False)


class GameSymmetries(object):
    """Namespace: one method per game name (the games compiled here)."""

    def breakthroughSmall(self):
        return Symmetries(skip_bases=["control"], apply_bases=[ApplySymmetry("cell", 1, 2)],
                          skip_actions=["noop"], apply_actions=[ApplySymmetry("move", [1, 3], [2, 4])],
                          do_reflection=True)

    def breakthrough(self):
        return Symmetries(skip_bases=["control"], apply_bases=[ApplySymmetry("cellHolds", 1, 2)],
                          skip_actions=["noop"], apply_actions=[ApplySymmetry("move", [1, 3], [2, 4])],
                          do_reflection=True)

    def reversi(self):
        return Symmetries(skip_bases=["control"], apply_bases=[ApplySymmetry("cell", 1, 2)],
                          skip_actions=["noop"], apply_actions=[ApplySymmetry("move", 1, 2)],
                          do_rotations_90=True, do_reflection=True)

    def connect6(self):
        return Symmetries(skip_bases=["control"], apply_bases=[ApplySymmetry("cell", 1, 2)],
                          skip_actions=["noop"], apply_actions=[ApplySymmetry("place", 1, 2)],
                          do_rotations_90=True, do_reflection=True)

    def _hex(self):
        return Symmetries(skip_bases=["control", "step", "owner", "canSwap"],
                          apply_bases=[ApplySymmetry("cell", 1, 2), ApplySymmetry("connected", 2, 3)],
                          skip_actions=["noop", "swap"], apply_actions=[ApplySymmetry("place", 1, 2)],
                          do_rotations_180=True)
    hexLG11 = _hex
    hexLG13 = _hex

    def amazons_10x10(self):
        return Symmetries(skip_bases=["turn"],
                          apply_bases=[ApplySymmetry("cell", 1, 2), ApplySymmetry("justMoved", 1, 2)],
                          skip_actions=["noop"],
                          apply_actions=[ApplySymmetry("move", [1, 3], [2, 4]), ApplySymmetry("fire", 1, 2)],
                          do_rotations_90=True, do_reflection=True)


def get_symmetries_desc(game):
    fn = getattr(GameSymmetries(), game, None)
    return fn() if fn else None
