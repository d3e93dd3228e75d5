"""KIF / GDL s-expression parser and term helpers.

Terms are represented as:
  * constants  -> interned ``str`` (e.g. ``'white'``, ``'1'``)
  * variables  -> interned ``str`` starting with ``'?'``
  * compounds  -> ``tuple`` whose first element is the functor (a constant)

Rules are ``Rule(head, body)``; body literals are terms whose functor may be
``not`` / ``distinct`` / ``or`` / ``true`` / ``does`` or a user relation.

The game descriptions consumed here are the GGP rulesheets the reference ships
under ``data/rulesheets/*.kif`` (the reference itself never parses them: ggplib
does, see ``doc/install.rst:11-18``).
"""
from __future__ import annotations

import sys
from typing import Iterable, List, NamedTuple, Tuple, Union

Term = Union[str, tuple]


class Rule(NamedTuple):
    head: Term
    body: Tuple[Term, ...]


def tokenize(text: str) -> List[str]:
    out: List[str] = []
    for line in text.splitlines():
        semi = line.find(';')
        if semi >= 0:
            line = line[:semi]
        out.extend(line.replace('(', ' ( ').replace(')', ' ) ').split())
    return out


def parse_sexprs(text: str) -> List[Term]:
    """Parse all top-level s-expressions. Symbol case is preserved (as ggplib does: the reference matches "cellHolds", "justMoved")."""
    toks = tokenize(text)
    pos = 0
    n = len(toks)
    intern = sys.intern

    def read() -> Term:
        nonlocal pos
        t = toks[pos]
        pos += 1
        if t == '(':
            items = []
            while toks[pos] != ')':
                items.append(read())
            pos += 1
            return tuple(items)
        if t == ')':
            raise ValueError("unbalanced ')'")
        return intern(t)

    res = []
    while pos < n:
        res.append(read())
    return res


def is_var(t: Term) -> bool:
    return type(t) is str and t.startswith('?')


def functor(t: Term) -> str:
    return t[0] if type(t) is tuple else t


def term_vars(t: Term, acc=None) -> list:
    """Variables of a term in first-occurrence order."""
    if acc is None:
        acc = []
    if type(t) is tuple:
        for a in t:
            term_vars(a, acc)
    elif t.startswith('?'):
        if t not in acc:
            acc.append(t)
    return acc


def is_ground(t: Term) -> bool:
    if type(t) is tuple:
        for a in t:
            if not is_ground(a):
                return False
        return True
    return not t.startswith('?')


def subst(t: Term, env: dict) -> Term:
    if type(t) is tuple:
        return tuple(subst(a, env) for a in t)
    if t.startswith('?'):
        v = env.get(t)
        if v is None:
            return t
        # bindings may themselves contain variables bound later
        return subst(v, env) if (type(v) is tuple or v.startswith('?')) else v
    return t


def term_to_str(t: Term) -> str:
    if type(t) is tuple:
        return '(' + ' '.join(term_to_str(a) for a in t) + ')'
    return t


def _expand_or(body: Iterable[Term]) -> List[Tuple[Term, ...]]:
    """Expand ``(or a b ...)`` literals into alternative conjunctive bodies (DNF)."""
    bodies: List[List[Term]] = [[]]
    for lit in body:
        if type(lit) is tuple and lit[0] == 'or':
            alts: List[Tuple[Term, ...]] = []
            for alt in lit[1:]:
                alts.extend(_expand_or([alt]))
            bodies = [b + list(a) for b in bodies for a in alts]
        elif type(lit) is tuple and lit[0] == 'and':
            subs = _expand_or(lit[1:])
            bodies = [b + list(a) for b in bodies for a in subs]
        else:
            for b in bodies:
                b.append(lit)
    return [tuple(b) for b in bodies]


def parse_rules(text: str) -> List[Rule]:
    """Parse a rulesheet into a list of rules (facts have an empty body); ``or`` is expanded."""
    rules: List[Rule] = []
    for sx in parse_sexprs(text):
        if type(sx) is tuple and sx and sx[0] == '<=':
            head = sx[1]
            for body in _expand_or(sx[2:]):
                rules.append(Rule(head, body))
        else:
            rules.append(Rule(sx, ()))
    return rules


def load_rules(path: str) -> List[Rule]:
    with open(path) as f:
        return parse_rules(f.read())
