"""Helpers to read tests/golden/*.json.gz fixtures and map their symbols onto a compiled propnet."""
import gzip
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_fixture(game):
    with gzip.open(os.path.join(GOLDEN, game + '.json.gz'), 'rt') as f:
        return json.load(f)


class FixtureMap:
    """Maps fixture base / input indices (by GDL string) to the indices of a compiled propnet."""

    def __init__(self, fx, pn):
        assert fx['roles'] == pn.roles
        # GDL symbols are case-insensitive: the interpreter that made the fixtures folds case, the
        # compiled tables keep the rulesheet's spelling (the reference's gamedesc uses it: kingHasMoved)
        b = {s.lower(): i for i, s in enumerate(pn.bases)}
        self.base = np.array([b[s.lower()] for s in fx['bases']], dtype=np.int64)
        assert len(set(self.base.tolist())) == len(pn.bases) == len(fx['bases'])
        self.inp = []
        for r in range(len(pn.roles)):
            m = {s.lower(): i for i, s in enumerate(pn.inputs[r])}
            # fixture inputs are the rulesheet's `input` relation; must be the same set
            assert set(x.lower() for x in fx['inputs'][r]) == set(x.lower() for x in pn.inputs[r])
            self.inp.append(np.array([m[s.lower()] for s in fx['inputs'][r]], dtype=np.int64))
        self.n_bases = len(pn.bases)

    def state_bits(self, idx_list):
        bits = np.zeros(self.n_bases, dtype=np.uint8)
        if idx_list:
            bits[self.base[np.array(idx_list, dtype=np.int64)]] = 1
        return bits

    def legal_set(self, role, idx_list):
        return sorted(int(self.inp[role][i]) for i in idx_list)

    def move(self, role, i):
        return int(self.inp[role][i])
