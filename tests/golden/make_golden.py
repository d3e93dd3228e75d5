#!/usr/bin/env python
"""Generates tests/golden/<game>.json.gz from the reference's rulesheets with the from-rules GDL
interpreter (oracle/gdl_interp.py). Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py [game:n_games:max_plies ...]

Each fixture holds random-play trajectories: per ply the full state, every role's legal move set,
terminal flag, goal values and the joint move played. Symbols are stored as GDL strings so that the
fixtures do not depend on any base/action index order.
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
DEFAULT = ['breakthrough:4:400', 'reversi:3:100', 'connect6:2:60', 'amazons_10x10:2:120',
           'breakthroughSmall:4:200', 'chess_15d:2:70', 'hexLG13:1:45']


def make(game, n_games, max_plies, seed=20260118):
    g = GdlInterpreter(open(os.path.join(RULES, game + '.kif')).read())
    bases = [to_str(b) for b in g.base_props()]
    bidx = {b: i for i, b in enumerate(bases)}
    inputs = {r: [to_str(m) for m in g.input_props(r)] for r in g.roles}
    iidx = {r: {m: i for i, m in enumerate(inputs[r])} for r in g.roles}
    rng = random.Random(seed)
    games = []
    for _ in range(n_games):
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
