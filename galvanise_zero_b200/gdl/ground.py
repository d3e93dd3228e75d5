"""GDL grounder: rules -> ground rule instances over base / input propositions.

Pipeline (all host side, run once per game, result cached as a compiled propnet table):

1. ``or`` expansion (parser), predicate classification (static / depends on ``true`` / ``does``).
2. **Existential factorisation**: body literals that are connected only through variables that do
   not occur in the head are split out into auxiliary predicates, so the number of ground instances
   is a sum over independent sub-conjunctions instead of their product (e.g. breakthrough's frame
   axiom ``next(cellHolds x3 y3 s) :- true(cellHolds x3 y3 s), does(p, move(x1 y1 x2 y2)),
   distinctCell(..)`` or hexLG13's ``connected`` transfer rule).
3. Static relations are evaluated exactly, bottom-up (semi-naive, stratified negation).
4. Dynamic relations: an over-approximating fixpoint in which ``true(X)`` ranges over all bases,
   ``does(r, m)`` over all inputs and dynamic negative literals are ignored.
5. A final pass enumerates every ground rule instance whose positive literals are possible; static
   literals are resolved and dropped. The instances feed ``propnet.build``.

Canonical ORDER of roles / bases / inputs (the reference inherits ggplib's order, which is not
recorded anywhere in the reference — SURVEY.md §8(c)): rule order in the rulesheet, then
nested-loop order over the rule's body literals as written, with every relation's facts in
derivation order (file order for plain facts). This reproduces the one ordering fact the reference
pins: reversi action 1 is ``(move 1 1)`` and action 8 is ``(move 1 8)``
(``src/test/test_symmetry.py:90-91``).
"""
from __future__ import annotations

import sys
from collections import defaultdict
from typing import Dict, List, Tuple

from .parser import Rule, Term, is_ground, subst, term_vars

STATIC, STATE, MOVE = 0, 1, 2


def _key(t: Term):
    return (t[0], len(t)) if type(t) is tuple else (t, 0)


def _atom(lit: Term) -> Term:
    while type(lit) is tuple and lit[0] == 'not':
        lit = lit[1]
    return lit


def _is_neg(lit: Term) -> bool:
    return type(lit) is tuple and lit[0] == 'not'


def _is_distinct(lit: Term) -> bool:
    return type(lit) is tuple and lit[0] == 'distinct'


# --------------------------------------------------------------------------------------------
class Relation:
    __slots__ = ('facts', 'fset', 'indexes')

    def __init__(self):
        self.facts: List[Term] = []
        self.fset = set()
        self.indexes: Dict[tuple, Dict[tuple, List[Term]]] = {}

    def add(self, fact: Term) -> bool:
        if fact in self.fset:
            return False
        self.fset.add(fact)
        self.facts.append(fact)
        for pos, idx in self.indexes.items():
            idx.setdefault(tuple(fact[p] for p in pos), []).append(fact)
        return True

    def lookup(self, pos: tuple, key: tuple):
        if not pos:
            return self.facts
        idx = self.indexes.get(pos)
        if idx is None:
            idx = {}
            for f in self.facts:
                idx.setdefault(tuple(f[p] for p in pos), []).append(f)
            self.indexes[pos] = idx
        return idx.get(key, ())


def _match(pat: Term, val: Term, env: dict, trail: list) -> bool:
    """One-way match of pattern (may contain variables) against a ground term."""
    if type(pat) is tuple:
        if type(val) is not tuple or len(pat) != len(val):
            return False
        for p, v in zip(pat, val):
            if not _match(p, v, env, trail):
                return False
        return True
    if pat.startswith('?'):
        b = env.get(pat)
        if b is None:
            env[pat] = val
            trail.append(pat)
            return True
        return b == val
    return pat == val


class _Plan:
    """A compiled join order for one rule (optionally with one literal restricted to a delta)."""
    __slots__ = ('steps', 'checks_at', 'head', 'pre')

    def __init__(self):
        self.steps = []      # (literal, relkey, bound_positions, free_positions)
        self.checks_at = []  # per step: list of (kind, literal) evaluated once bound
        self.head = None


class Grounder:
    def __init__(self, rules: List[Rule], verbose: bool = False):
        self.verbose = verbose
        self.rules = [Rule(r.head, tuple(r.body)) for r in rules]
        self._aux_count = 0
        self._classify()
        self._factorise()
        self._classify()
        self.rel: Dict[tuple, Relation] = defaultdict(Relation)

    # ------------------------------------------------------------------ classification
    def _classify(self):
        by_head = defaultdict(list)
        for r in self.rules:
            by_head[_key(r.head)].append(r)
        self.by_head = by_head
        deps = defaultdict(set)
        for r in self.rules:
            k = _key(r.head)
            for lit in r.body:
                if _is_distinct(lit):
                    continue
                deps[k].add(_key(_atom(lit)))
        cls: Dict[tuple, int] = {('true', 2): STATE, ('does', 3): MOVE}
        for k in by_head:
            cls.setdefault(k, STATIC)
        changed = True
        while changed:
            changed = False
            for k, ds in deps.items():
                c = cls.get(k, STATIC)
                for d in ds:
                    dc = cls.get(d, STATIC)
                    if dc > c:
                        c = dc
                if c != cls.get(k, STATIC):
                    cls[k] = c
                    changed = True
        self.cls = cls

    def _lit_class(self, lit: Term) -> int:
        if _is_distinct(lit):
            return STATIC
        return self.cls.get(_key(_atom(lit)), STATIC)

    # ------------------------------------------------------------------ factorisation
    def _factorise(self):
        """Split bodies into components connected through existential (non-head) variables."""
        out: List[Rule] = []
        work = list(self.rules)
        while work:
            r = work.pop(0)
            if len(r.body) < 2 or self.cls.get(_key(r.head), STATIC) == STATIC:
                out.append(r)
                continue
            hv = set(term_vars(r.head))
            lit_ex = [set(v for v in term_vars(l) if v not in hv) for l in r.body]
            n = len(r.body)
            parent = list(range(n))

            def find(i):
                while parent[i] != i:
                    parent[i] = parent[parent[i]]
                    i = parent[i]
                return i

            for i in range(n):
                for j in range(i + 1, n):
                    if lit_ex[i] & lit_ex[j]:
                        parent[find(i)] = find(j)
            comps = defaultdict(list)
            for i in range(n):
                comps[find(i)].append(i)
            groups = [g for g in comps.values() if any(lit_ex[i] for i in g)]
            main = [i for g in comps.values() if not any(lit_ex[i] for i in g) for i in g]
            # only worth splitting if there are >= 2 independent parts
            if len(groups) + (1 if main else 0) < 2:
                out.append(r)
                continue
            # literals without existential vars that only *test* (not / distinct) need their
            # variables bound; they stay in the main rule where aux heads bind the head vars.
            new_body: List[Tuple[int, Term]] = [(i, r.body[i]) for i in main]
            for g in groups:
                lits = [r.body[i] for i in sorted(g)]
                # a component must be able to bind all its variables positively on its own
                pos_vars = set()
                for l in lits:
                    if not _is_neg(l) and not _is_distinct(l):
                        pos_vars.update(term_vars(l))
                all_vars = set()
                for l in lits:
                    all_vars.update(term_vars(l))
                if not all_vars <= pos_vars:
                    # negative literal needs head vars bound elsewhere: keep it un-split
                    new_body = None
                    break
                gv = [v for v in term_vars(tuple(lits)) if v in hv]
                self._aux_count += 1
                name = sys.intern('__aux%d' % self._aux_count)
                aux_head = (name,) + tuple(gv) if gv else name
                aux_rule = Rule(aux_head, tuple(lits))
                # classify new predicate lazily: assume dynamic if any literal is
                c = max(self._lit_class(l) for l in lits)
                self.cls[_key(aux_head)] = c
                work.append(aux_rule) if c != STATIC else out.append(aux_rule)
                new_body.append((min(g), aux_head))
            if new_body is None:
                out.append(r)
                continue
            new_body.sort(key=lambda t: t[0])
            out.append(Rule(r.head, tuple(l for _, l in new_body)))
        self.rules = out

    # ------------------------------------------------------------------ join engine
    def _plan(self, rule: Rule, first: int = -1, ignore_dyn_neg: bool = True,
              keep_order: bool = False) -> _Plan:
        """Greedy join order. ``first``: index of the literal to evaluate first (delta literal)."""
        body = list(rule.body)
        pos_idx = [i for i, l in enumerate(body) if not _is_neg(l) and not _is_distinct(l)]
        chk_idx = [i for i, l in enumerate(body) if _is_neg(l) or _is_distinct(l)]
        bound: set = set()
        order = []
        remaining = list(pos_idx)
        if first >= 0:
            order.append(first)
            remaining.remove(first)
            bound.update(term_vars(body[first]))
        while remaining:
            if keep_order:
                best = remaining[0]
            else:
                best, best_score = None, None
                for i in remaining:
                    lit = body[i]
                    args = lit[1:] if type(lit) is tuple else ()
                    nb = sum(1 for a in args if all(v in bound for v in term_vars(a)))
                    nfree = len(args) - nb
                    size = len(self.rel[_key(lit)].facts) if _key(lit) in self.rel else 1 << 30
                    # prefer fully bound, then many bound args, then small relations
                    score = (0 if nfree == 0 else 1, -nb, size, i)
                    if best_score is None or score < best_score:
                        best, best_score = i, score
            order.append(best)
            remaining.remove(best)
            bound.update(term_vars(body[best]))
        plan = _Plan()
        plan.head = rule.head
        bound = set()
        pending = list(chk_idx)

        def classify_check(cl):
            if _is_distinct(cl):
                return ('distinct', cl)
            a = _atom(cl)
            if self._lit_class(cl) == STATIC:
                return ('notstatic', a)
            return ('notdyn', a)

        plan.pre = []
        for c in list(pending):
            if is_ground(body[c]):
                pending.remove(c)
                plan.pre.append(classify_check(body[c]))
        for i in order:
            lit = body[i]
            args = lit[1:] if type(lit) is tuple else ()
            bpos = tuple(p + 1 for p, a in enumerate(args) if all(v in bound for v in term_vars(a)))
            plan.steps.append((lit, _key(lit), bpos))
            bound.update(term_vars(lit))
            now = []
            for c in list(pending):
                cl = body[c]
                if all(v in bound for v in term_vars(cl)):
                    pending.remove(c)
                    now.append(classify_check(cl))
            plan.checks_at.append(now)
        if pending:
            # checks with no positive literal at all (fully ground literals)
            raise ValueError('unsafe rule (unbound variable in not/distinct): %r' % (rule,))
        return plan

    def _run(self, plan: _Plan, emit, delta: List[Term] = None):
        """Enumerate all bindings of ``plan``; calls emit(env, dyn_negs)."""
        steps = plan.steps
        checks = plan.checks_at
        nsteps = len(steps)
        rel = self.rel
        env: dict = {}

        def rec(si: int, negs: tuple):
            if si == nsteps:
                emit(env, negs)
                return
            lit, rk, bpos = steps[si]
            if si == 0 and delta is not None:
                cands = delta
                if bpos:  # cannot happen for the first literal, but be safe
                    key = tuple(subst(lit[p], env) for p in bpos)
                    cands = [f for f in delta if tuple(f[p] for p in bpos) == key]
            else:
                r = rel.get(rk)
                if r is None:
                    return
                if bpos:
                    key = tuple(subst(lit[p], env) for p in bpos)
                    cands = r.lookup(bpos, key)
                else:
                    cands = r.facts
                if len(bpos) == (len(lit) - 1 if type(lit) is tuple else 0):
                    # fully bound: membership test
                    g = subst(lit, env) if type(lit) is tuple else lit
                    cands = (g,) if g in r.fset else ()
            cl = checks[si]
            for f in cands:
                trail: list = []
                if type(lit) is tuple:
                    ok = _match(lit, f, env, trail)
                else:
                    ok = (lit == f)
                if ok:
                    n2 = negs
                    for kind, c in cl:
                        if kind == 'distinct':
                            if subst(c[1], env) == subst(c[2], env):
                                ok = False
                                break
                        elif kind == 'notstatic':
                            g = subst(c, env)
                            r2 = rel.get(_key(g))
                            if r2 is not None and g in r2.fset:
                                ok = False
                                break
                        else:
                            n2 = n2 + (subst(c, env),)
                    if ok:
                        rec(si + 1, n2)
                for v in trail:
                    del env[v]

        negs0 = ()
        for kind, c in plan.pre:
            if kind == 'distinct':
                if c[1] == c[2]:
                    return
            elif kind == 'notstatic':
                r2 = rel.get(_key(c))
                if r2 is not None and c in r2.fset:
                    return
            else:
                negs0 = negs0 + (c,)
        if nsteps == 0:
            emit(env, negs0)
            return
        rec(0, negs0)

    # ------------------------------------------------------------------ evaluation
    def _eval_stratum(self, rules: List[Rule], growing: set):
        """Semi-naive fixpoint for ``rules``; ``growing`` = relation keys that may gain facts."""
        rel = self.rel
        new_facts: Dict[tuple, List[Term]] = defaultdict(list)

        def make_emit(head):
            hk = _key(head)
            hrel = rel[hk]
            ground_head = is_ground(head)

            def emit(env, negs):
                h = head if ground_head else subst(head, env)
                if hrel.add(h):
                    new_facts[hk].append(h)
            return emit

        # round 0: everything
        for r in rules:
            if not r.body:
                if rel[_key(r.head)].add(r.head):
                    new_facts[_key(r.head)].append(r.head)
                continue
            self._run(self._plan(r), make_emit(r.head))
        rounds = 0
        while new_facts:
            rounds += 1
            delta, new_facts = new_facts, defaultdict(list)
            for r in rules:
                for i, lit in enumerate(r.body):
                    if _is_neg(lit) or _is_distinct(lit):
                        continue
                    k = _key(lit)
                    if k in growing and k in delta:
                        self._run(self._plan(r, first=i), make_emit(r.head), delta=delta[k])
        return rounds

    def ground(self):
        """Returns dict with roles, bases, inputs, init and ground instances."""
        rel = self.rel
        cls = self.cls
        static_rules = [r for r in self.rules if cls.get(_key(r.head), STATIC) == STATIC]
        dyn_rules = [r for r in self.rules if cls.get(_key(r.head), STATIC) != STATIC]

        # ---- static part, stratified by negation (simple iterative stratification)
        self._eval_static(static_rules)

        # ---- canonical orderings from role / base / input / init
        roles = [f[1] for f in rel[('role', 2)].facts]
        bases = self._ordered('base', 2, lambda f: f[1])
        inputs = {r: [] for r in roles}
        for f in self._ordered('input', 3, lambda f: f):
            inputs[f[1]].append(f[2])
        init = [f[1] for f in rel[('init', 2)].facts]

        # ---- dynamic over-approximation
        for b in bases:
            rel[('true', 2)].add(('true', b))
        for r in roles:
            for m in inputs[r]:
                rel[('does', 3)].add(('does', r, m))
        growing = set(_key(r.head) for r in dyn_rules)
        rounds = self._eval_stratum(dyn_rules, growing)
        if self.verbose:
            tot = sum(len(rel[k].facts) for k in growing)
            print('[ground] dynamic fixpoint: %d rounds, %d possible atoms' % (rounds, tot),
                  file=sys.stderr)

        # ---- instance enumeration
        instances: Dict[Term, List[Tuple[tuple, tuple]]] = defaultdict(list)
        ninst = 0
        for r in dyn_rules:
            head = r.head
            dyn_pos = [l for l in r.body if not _is_neg(l) and not _is_distinct(l)
                       and self._lit_class(l) != STATIC]

            def emit(env, negs, head=head, dyn_pos=dyn_pos):
                nonlocal ninst
                h = subst(head, env)
                pos = tuple(subst(l, env) for l in dyn_pos)
                # negative dynamic literals whose atom is impossible are always satisfied
                neg = tuple(a for a in negs if a in rel[_key(a)].fset)
                instances[h].append((pos, neg))
                ninst += 1

            if not r.body:
                instances[head].append(((), ()))
                continue
            self._run(self._plan(r), emit)
        if self.verbose:
            print('[ground] %d heads, %d instances' % (len(instances), ninst), file=sys.stderr)
        return {
            'roles': roles, 'bases': bases, 'inputs': inputs, 'init': init,
            'instances': instances,
        }

    def _ordered(self, name: str, arity: int, proj):
        """Facts of a static relation in canonical order (rule order, literal nested-loop order)."""
        seen = set()
        out = []
        for r in self.by_head.get((name, arity), []):
            if not r.body:
                if r.head not in seen:
                    seen.add(r.head)
                    out.append(proj(r.head))
                continue

            def emit(env, negs, r=r):
                h = subst(r.head, env)
                if h not in seen:
                    seen.add(h)
                    out.append(proj(h))
            self._run(self._plan(r, keep_order=True), emit)
        return out

    def _eval_static(self, rules: List[Rule]):
        """Stratify by negation and evaluate each stratum to fixpoint."""
        heads = set(_key(r.head) for r in rules)
        stratum = {k: 0 for k in heads}
        changed = True
        guard = 0
        while changed:
            changed = False
            guard += 1
            if guard > len(heads) + 5:
                raise ValueError('static rules are not stratifiable')
            for r in rules:
                hk = _key(r.head)
                for lit in r.body:
                    if _is_distinct(lit):
                        continue
                    ak = _key(_atom(lit))
                    if ak not in heads:
                        continue
                    need = stratum[ak] + (1 if _is_neg(lit) else 0)
                    if need > stratum[hk]:
                        stratum[hk] = need
                        changed = True
        for s in sorted(set(stratum.values())):
            rs = [r for r in rules if stratum[_key(r.head)] == s]
            self._eval_stratum(rs, set(_key(r.head) for r in rs))
