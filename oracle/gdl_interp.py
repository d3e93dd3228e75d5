"""ORACLE (test infrastructure, never imported by the product path).

From-rules GDL interpreter: answers ``legal`` / ``next`` / ``terminal`` / ``goal`` queries directly
from a rulesheet by tabled top-down resolution with negation-as-failure. It never builds a propnet,
so it is an *independent* semantic check of the propnet compiler in ``galvanise_zero_b200/gdl`` and
of every state machine built on its tables.

What it restates: the game semantics that the reference obtains from ggplib's state machine
(``StateMachineInterface::getLegalState / nextState / isTerminal / getGoalValue``, called at
``src/cpp/puct/node.cpp:115-120,157-217`` and ``src/cpp/puct/evaluator.cpp:218-219``). ggplib
(github richemslie/ggplib, no version pinned by the reference: ``src/cpp/Makefile:1-7``,
``bin/setup.sh``) is not vendored and not installed, so the published GDL semantics
(Love et al., "General Game Playing: Game Description Language Specification", 2008) over the
reference's own rulesheets ``data/rulesheets/*.kif`` is the ground truth here.
PARITY STATUS: semantic (set-level) oracle; index ORDER of bases/actions is unpinned by the reference.

Pure Python; intended for small fixtures only (milliseconds to seconds per state).
"""
from __future__ import annotations

import sys
from collections import defaultdict


# ----------------------------------------------------------------------------- tiny parser
def _tokenize(text):
    out = []
    for line in text.splitlines():
        i = line.find(';')
        if i >= 0:
            line = line[:i]
        out.extend(line.replace('(', ' ( ').replace(')', ' ) ').split())
    return out


def parse(text):
    toks = _tokenize(text)
    pos = 0

    def read():
        nonlocal pos
        t = toks[pos]
        pos += 1
        if t == '(':
            items = []
            while toks[pos] != ')':
                items.append(read())
            pos += 1
            return tuple(items)
        return sys.intern(t)

    res = []
    while pos < len(toks):
        res.append(read())
    return res


def to_str(t):
    if type(t) is tuple:
        return '(' + ' '.join(to_str(a) for a in t) + ')'
    return t


# ----------------------------------------------------------------------------- unification
def _walk(t, env):
    while type(t) is str and t.startswith('?'):
        v = env.get(t)
        if v is None:
            return t
        t = v
    return t


def _resolve(t, env):
    t = _walk(t, env)
    if type(t) is tuple:
        return tuple(_resolve(a, env) for a in t)
    return t


def _unify(a, b, env):
    """Returns an extended env (new dict when extended) or None."""
    a = _walk(a, env)
    b = _walk(b, env)
    if a is b or a == b:
        return env
    if type(a) is str and a.startswith('?'):
        env = dict(env)
        env[a] = b
        return env
    if type(b) is str and b.startswith('?'):
        env = dict(env)
        env[b] = a
        return env
    if type(a) is tuple and type(b) is tuple and len(a) == len(b):
        for x, y in zip(a, b):
            env = _unify(x, y, env)
            if env is None:
                return None
        return env
    return None


def _ground(t):
    if type(t) is tuple:
        return all(_ground(a) for a in t)
    return not t.startswith('?')


def _rename(t, suffix):
    if type(t) is tuple:
        return tuple(_rename(a, suffix) for a in t)
    if t.startswith('?'):
        return sys.intern(t + suffix)
    return t


def _canon(t, mapping):
    if type(t) is tuple:
        return tuple(_canon(a, mapping) for a in t)
    if t.startswith('?'):
        m = mapping.get(t)
        if m is None:
            m = mapping[t] = '?%d' % len(mapping)
        return m
    return t


def _key(t):
    return (t[0], len(t)) if type(t) is tuple else (t, 0)


_STATIC, _STATE, _MOVE = 0, 1, 2


class GdlInterpreter:
    def __init__(self, text):
        self.rules = defaultdict(list)         # (functor, arity) -> [(head, body)]
        order = []
        for sx in parse(text):
            if type(sx) is tuple and sx and sx[0] == '<=':
                head, bodies = sx[1], self._expand_or(list(sx[2:]))
            else:
                head, bodies = sx, [[]]
            for body in bodies:
                k = _key(head)
                if k not in self.rules:
                    order.append(k)
                self.rules[k].append((head, tuple(body)))
        self._classify()
        self._fresh = 0
        self._tables = ({}, {}, {})
        self._in_progress = set()
        self.state = frozenset()
        self._state_by_key = {}
        self.moves = {}
        self.roles = [a[1] for a in self.ask(('role', '?r'))]

    # -- preprocessing
    def _expand_or(self, body):
        bodies = [[]]
        for lit in body:
            if type(lit) is tuple and lit[0] == 'or':
                alts = []
                for alt in lit[1:]:
                    alts.extend(self._expand_or([alt]))
                bodies = [b + a for b in bodies for a in alts]
            else:
                bodies = [b + [lit] for b in bodies]
        return bodies

    def _classify(self):
        """Dependency class per predicate: static / depends on true / depends on does."""
        deps = defaultdict(set)
        for k, rs in self.rules.items():
            for _, body in rs:
                for lit in body:
                    while type(lit) is tuple and lit[0] == 'not':
                        lit = lit[1]
                    deps[k].add(_key(lit))
        cls = {}

        def visit(k, stack):
            if k in cls:
                return cls[k]
            if k[0] == 'true':
                return _STATE
            if k[0] == 'does':
                return _MOVE
            if k in stack:
                return _STATIC
            stack.add(k)
            c = _STATIC
            for d in deps.get(k, ()):
                c = max(c, visit(d, stack))
            stack.discard(k)
            return c

        # iterate to a fixpoint so that mutually recursive predicates settle
        for _ in range(3):
            cls = {}
            for k in list(self.rules):
                cls[k] = visit(k, set())
        self.cls = cls

    # -- state
    def set_state(self, facts):
        self.state = frozenset(facts)
        by = defaultdict(list)
        for f in self.state:
            by[_key(f)].append(f)
        self._state_by_key = by
        self._tables[_STATE].clear()
        self._tables[_MOVE].clear()

    def set_moves(self, moves):
        """moves: dict role -> move term (or None to clear)."""
        self.moves = dict(moves or {})
        self._tables[_MOVE].clear()

    # -- resolution
    def answers(self, pattern):
        """All ground instances of ``pattern`` (an atom, possibly with variables) that hold."""
        f = pattern[0] if type(pattern) is tuple else pattern
        if f == 'true':
            arg = pattern[1]
            if _ground(arg):
                return [pattern] if arg in self.state else []
            res = []
            for fact in self._state_by_key.get(_key(arg), ()):
                if _unify(arg, fact, {}) is not None:
                    res.append(('true', fact))
            return res
        if f == 'does':
            res = []
            for r, m in self.moves.items():
                if _unify(pattern, ('does', r, m), {}) is not None:
                    res.append(('does', r, m))
            return res
        k = _key(pattern)
        rs = self.rules.get(k)
        if not rs:
            return []
        table = self._tables[self.cls.get(k, _STATIC)]
        ck = _canon(pattern, {})
        got = table.get(ck)
        if got is not None:
            return got
        if ck in self._in_progress:
            raise RuntimeError('cyclic call pattern: ' + to_str(ck))
        self._in_progress.add(ck)
        seen = set()
        out = []
        for head, body in rs:
            self._fresh += 1
            suffix = '_%d' % self._fresh
            h = _rename(head, suffix)
            env = _unify(h, pattern, {})
            if env is None:
                continue
            b = [_rename(l, suffix) for l in body]
            for e in self._solve(b, env):
                ans = _resolve(pattern, e)
                if not _ground(ans):
                    raise RuntimeError('unsafe rule, non-ground answer: ' + to_str(ans))
                if ans not in seen:
                    seen.add(ans)
                    out.append(ans)
        self._in_progress.discard(ck)
        table[ck] = out
        return out

    def _solve(self, lits, env):
        if not lits:
            yield env
            return
        # pick the first literal that is evaluable now (defer non-ground not/distinct)
        idx = 0
        for idx, lit in enumerate(lits):
            if type(lit) is tuple and lit[0] in ('not', 'distinct'):
                if _ground(_resolve(lit, env)):
                    break
            else:
                break
        else:
            raise RuntimeError('floundering: ' + ' '.join(to_str(_resolve(l, env)) for l in lits))
        lit = lits[idx]
        rest = lits[:idx] + lits[idx + 1:]
        if type(lit) is tuple and lit[0] == 'distinct':
            a, b = _resolve(lit[1], env), _resolve(lit[2], env)
            if a != b:
                yield from self._solve(rest, env)
            return
        if type(lit) is tuple and lit[0] == 'not':
            inner = _resolve(lit[1], env)
            if not self.answers(inner):
                yield from self._solve(rest, env)
            return
        goal = _resolve(lit, env)
        for ans in self.answers(goal):
            e = _unify(goal, ans, env)
            if e is not None:
                yield from self._solve(rest, e)

    def ask(self, pattern):
        return self.answers(pattern)

    # -- state-machine style API
    def initial_state(self):
        return frozenset(a[1] for a in self.ask(('init', '?x')))

    def base_props(self):
        return [a[1] for a in self.ask(('base', '?x'))]

    def input_props(self, role):
        return [a[2] for a in self.ask(('input', role, '?m'))]

    def legal_moves(self, role):
        return [a[2] for a in self.ask(('legal', role, '?m'))]

    def is_terminal(self):
        return bool(self.ask('terminal'))

    def goal(self, role):
        vals = [int(a[2]) for a in self.ask(('goal', role, '?v'))]
        return vals

    def next_state(self, moves):
        self.set_moves(moves)
        res = frozenset(a[1] for a in self.ask(('next', '?x')))
        self.set_moves(None)
        return res
