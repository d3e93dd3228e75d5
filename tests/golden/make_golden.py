#!/usr/bin/env python
"""Generates tests/golden/<game>.json.gz from the reference's rulesheets with the from-rules GDL
interpreter (oracle/gdl_interp.py). Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py [game:n_games:max_plies ...]

Each fixture holds random-play trajectories: per ply the full state, every role's legal move set,
terminal flag, goal values and the joint move played. Symbols are stored as GDL strings so that the
fixtures do not depend on any base/action index order.

Games are played to their end (up to max_plies). Some games of a fixture are STEERED towards the rules'
corner cases — the moves are still chosen among the interpreter's legal moves, a preference only biases the
choice — so that the fixtures pin them (tests/test_oracle_propnet.py::test_fixtures_cover_the_rules_corner_cases):
  reversi     forced passes and games ending on a double pass / full board (they occur in plain random play)
  hexLG13     white answers black's first stone with `swap` (every second game)
  chess_15d   castling on both wings (scripted openings), an en-passant capture (scripted), promotions
              (pawn-rush games), the 32-step counter running out (quiet-move games), check evasions (natural)
"""
import gzip
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
from oracle.gdl_interp import GdlInterpreter, to_str  # noqa: E402

RULES = '/root/reference/data/rulesheets'
DEFAULT = ['breakthrough:20:400', 'reversi:20:100', 'connect6:20:90', 'amazons_10x10:20:120',
           'breakthroughSmall:20:200', 'chess_15d:20:220', 'hexLG13:6:175']

# chess: scripted preferences (taken when legal, else a random legal move), one list per role
CHESS_SCRIPTS = [
    # both sides castle kingside
    {'white': ['(move knight g 1 f 3)', '(move pawn e 2 e 3)', '(move bishop f 1 e 2)', 'castleKingside'],
     'black': ['(move knight g 8 f 6)', '(move pawn e 7 e 6)', '(move bishop f 8 e 7)', 'castleKingside']},
    # both sides castle queenside
    {'white': ['(move knight b 1 c 3)', '(move pawn d 2 d 3)', '(move bishop c 1 e 3)', '(move queen d 1 d 2)',
               'castleQueenside'],
     'black': ['(move knight b 8 c 6)', '(move pawn d 7 d 6)', '(move bishop c 8 e 6)', '(move queen d 8 d 7)',
               'castleQueenside']},
    # en passant: e2-e4, e4-e5, ... d7-d5, e5xd6
    {'white': ['(move pawn e 2 e 4)', '(move pawn e 4 e 5)', '(move pawn e 5 d 6)'],
     'black': ['(move pawn a 7 a 6)', '(move pawn d 7 d 5)']},
    # en passant by black: ... d7-d5, d5-d4, e2-e4, d4xe3
    {'white': ['(move pawn a 2 a 3)', '(move pawn h 2 h 3)', '(move pawn e 2 e 4)'],
     'black': ['(move pawn d 7 d 5)', '(move pawn d 5 d 4)', '(move pawn d 4 e 3)']},
]


def preference(game, game_index, ply, role, legal, rng, memo):
    """A legal move string to prefer, or None (uniformly random)."""
    if game == 'hexLG13':
        if game_index % 2 == 1 and 'swap' in legal:
            return 'swap'
        return None
    if game != 'chess_15d':
        return None
    if game_index < len(CHESS_SCRIPTS):
        script = memo.setdefault((game_index, role), list(CHESS_SCRIPTS[game_index].get(role, [])))
        if len(legal) == 1:
            return None              # not this role's turn (noop)
        while script:
            if script[0] in legal:
                return script.pop(0)
            script.pop(0)            # not possible (any more): give up on this step
        return None
    mode = game_index % 4
    if mode == 0 and rng.random() < 0.85:          # pawn rush: the most advanced pawn move available
        pawns = [m for m in legal if m.startswith('(move pawn') or m.split()[-1] in ('8)', '1)')]
        if pawns:
            key = (lambda m: int(m.split()[-1].rstrip(')'))) if role == 'white' else (lambda m: -int(m.split()[-1].rstrip(')')))
            return max(pawns, key=key)
    if mode == 1 and rng.random() < 0.9:           # quiet play: no pawn moves, so that the step counter runs
        quiet = [m for m in legal if m.startswith('(move knight') or m.startswith('(move rook') or m.startswith('(move king')]
        if quiet:
            return rng.choice(quiet)
    return None


def make(game, n_games, max_plies, seed=20260118):
    g = GdlInterpreter(open(os.path.join(RULES, game + '.kif')).read())
    bases = [to_str(b) for b in g.base_props()]
    bidx = {b: i for i, b in enumerate(bases)}
    inputs = {r: [to_str(m) for m in g.input_props(r)] for r in g.roles}
    iidx = {r: {m: i for i, m in enumerate(inputs[r])} for r in g.roles}
    rng = random.Random(seed)
    games = []
    memo = {}
    for gi in range(n_games):
        s = g.initial_state()
        plies = []
        for ply in range(max_plies):
            g.set_state(s)
            term = g.is_terminal()
            legal, moves = {}, {}
            for r in g.roles:
                lm = g.legal_moves(r)
                legal[r] = sorted(iidx[r][to_str(m)] for m in lm)
                moves[r] = rng.choice(lm) if lm else None
                if lm and not term:
                    strs = {to_str(m): m for m in lm}
                    want = preference(game, gi, ply, r, strs, rng, memo)
                    if want is not None and want in strs:
                        moves[r] = strs[want]
            rec = {
                'state': sorted(bidx[to_str(f)] for f in s),
                'legal': [legal[r] for r in g.roles],
                'terminal': int(term),
                'goals': [sorted(g.goal(r)) for r in g.roles],
            }
            if term or ply == max_plies - 1:
                rec['move'] = None
                plies.append(rec)
                break
            rec['move'] = [iidx[r][to_str(moves[r])] for r in g.roles]
            plies.append(rec)
            s = g.next_state(moves)
        games.append(plies)
    return {'game': game, 'roles': g.roles, 'bases': bases, 'inputs': [inputs[r] for r in g.roles],
            'games': games, 'generator': 'tests/golden/make_golden.py (oracle/gdl_interp.py)'}


if __name__ == '__main__':
    specs = sys.argv[1:] or DEFAULT
    for spec in specs:
        game, n, p = spec.split(':')
        fx = make(game, int(n), int(p))
        out = os.path.join(HERE, game + '.json.gz')
        with gzip.open(out, 'wt') as f:
            json.dump(fx, f, separators=(',', ':'))
        print(game, 'games', len(fx['games']), 'plies', sum(len(x) for x in fx['games']),
              '%.1f KB' % (os.path.getsize(out) / 1024))
