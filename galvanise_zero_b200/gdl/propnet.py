"""Propnet construction, optimisation and lowering to levelised AND-gate programs.

Normal form: every gate is an n-ary AND over *literals*; a literal is ``slot*2 + inverted``.
OR is stored through De Morgan (an inverted reference to an AND of inverted literals), so the
evaluators (``oracle/propnet_ref.c``, ``csrc/propnet.cuh``) need a single opcode:

    val[out] = AND_k ( val[lit_k >> 1] XOR -(lit_k & 1) )

Slot map:  0 = constant TRUE | 1..B = bases | B+1..B+I = inputs (role-major) | then gates.

The role that ggplib's propnet plays for the reference (``StateMachineInterface`` used at
``src/cpp/puct/node.cpp:153-221`` and ``src/cpp/puct/evaluator.cpp:216-239``) is played here by
three programs over one slot space:

* ``V``   — cone of legal / terminal / goal outputs (depends on bases only),
* ``T``   — cone of the next-state outputs (depends on bases and inputs),
* ``TV``  — ``T`` minus the gates already evaluated by ``V``;
* ``LT`` / ``GL`` — ``V`` split into the legal+terminal cone and the remaining goal-only gates;
* ``TL``  — ``T`` minus ``LT`` (next-state cone when only ``LT`` is known to be current).

One soundness assumption is used while optimising: **exactly one input per role is true** — which
holds by construction, because moves cross the C-ABI as one action index per role.
"""
from __future__ import annotations

import sys
from typing import Dict, List, Tuple

import numpy as np

TRUE = 0
FALSE = 1


def neg(l: int) -> int:
    return l ^ 1


class NetBuilder:
    def __init__(self, n_bases: int, role_sizes: List[int]):
        self.n_bases = n_bases
        self.role_sizes = list(role_sizes)
        self.n_inputs = sum(role_sizes)
        self.first_gate = 1 + n_bases + self.n_inputs
        self.gate_lits: List[Tuple[int, ...]] = []   # node id = first_gate + index
        self.cons: Dict[Tuple[int, ...], int] = {}
        # input node -> (role, action)
        self.input_role = {}
        node = 1 + n_bases
        self.input_base = []
        for r, n in enumerate(role_sizes):
            self.input_base.append(node)
            for a in range(n):
                self.input_role[node + a] = r
            node += n

    def base_lit(self, b: int) -> int:
        return (1 + b) * 2

    def input_lit(self, r: int, a: int) -> int:
        return (self.input_base[r] + a) * 2

    def lits_of(self, node: int) -> Tuple[int, ...]:
        return self.gate_lits[node - self.first_gate]

    def is_gate(self, node: int) -> bool:
        return node >= self.first_gate

    def mk_and(self, lits) -> int:
        s = set()
        for l in lits:
            if l == TRUE:
                continue
            if l == FALSE:
                return FALSE
            if (l ^ 1) in s:
                return FALSE
            s.add(l)
        if not s:
            return TRUE
        # exactly-one-input-per-role reasoning
        pos_by_role = {}
        neg_by_role = {}
        for l in s:
            r = self.input_role.get(l >> 1)
            if r is not None:
                (neg_by_role if (l & 1) else pos_by_role).setdefault(r, []).append(l)
        for r, ps in pos_by_role.items():
            if len(ps) > 1:
                return FALSE                       # two different moves of one role
            for l in neg_by_role.get(r, ()):       # implied by the positive literal
                s.discard(l)
            neg_by_role.pop(r, None)
        for r, ns in neg_by_role.items():
            n_r = self.role_sizes[r]
            if len(ns) * 2 > n_r:
                # AND_{m in S} !does(m)  ==  !AND_{m not in S} !does(m)
                inside = set(l >> 1 for l in ns)
                base = self.input_base[r]
                rest = [((base + a) * 2) | 1 for a in range(n_r) if (base + a) not in inside]
                for l in ns:
                    s.discard(l)
                c = neg(self.mk_and(rest))
                if c == FALSE:
                    return FALSE
                if c != TRUE:
                    if (c ^ 1) in s:
                        return FALSE
                    s.add(c)
        if not s:
            return TRUE
        if len(s) == 1:
            return next(iter(s))
        key = tuple(sorted(s))
        n = self.cons.get(key)
        if n is None:
            n = self.first_gate + len(self.gate_lits)
            self.gate_lits.append(key)
            self.cons[key] = n
        return n * 2

    def mk_or(self, lits) -> int:
        return neg(self.mk_and([neg(l) for l in lits]))


class Propnet:
    """Optimised, lowered propnet: everything the evaluators need, as numpy arrays."""

    MAX_FANIN = 8

    def __init__(self):
        self.roles: List[str] = []
        self.bases: List[str] = []          # printable GDL terms, canonical order
        self.inputs: List[List[str]] = []   # per role
        self.init_state: np.ndarray = None  # uint8 [n_bases]
        self.n_slots = 0
        self.gate_out: np.ndarray = None    # int32 [n_gates] slot written
        self.gate_lits: List[np.ndarray] = []
        self.gate_level: np.ndarray = None
        self.next_lit: np.ndarray = None    # int32 [n_bases]
        self.legal_lit: List[np.ndarray] = []  # per role int32 [n_actions]
        self.terminal_lit: int = FALSE
        self.goal_values: List[np.ndarray] = []  # per role int32 [n_goal]
        self.goal_lit: List[np.ndarray] = []     # per role int32 [n_goal]
        self.prog: Dict[str, np.ndarray] = {}    # 'V','T','TV' -> gate indices in level order

    @property
    def n_bases(self):
        return len(self.bases)

    @property
    def role_sizes(self):
        return [len(x) for x in self.inputs]


def build_propnet(g: dict, verbose: bool = False) -> Propnet:
    """``g`` is the dict produced by ``Grounder.ground()``."""
    from .parser import term_to_str

    sys.setrecursionlimit(max(sys.getrecursionlimit(), 200000))
    roles = g['roles']
    bases = g['bases']
    inputs = g['inputs']
    instances = g['instances']
    base_index = {b: i for i, b in enumerate(bases)}
    role_index = {r: i for i, r in enumerate(roles)}
    input_index = [{m: i for i, m in enumerate(inputs[r])} for r in roles]
    nb = NetBuilder(len(bases), [len(inputs[r]) for r in roles])

    memo: Dict[object, int] = {}
    in_progress = set()

    def atom_lit(a) -> int:
        f = a[0] if type(a) is tuple else a
        if f == 'true':
            i = base_index.get(a[1])
            return FALSE if i is None else nb.base_lit(i)
        if f == 'does':
            ri = role_index.get(a[1])
            if ri is None:
                return FALSE
            ai = input_index[ri].get(a[2])
            return FALSE if ai is None else nb.input_lit(ri, ai)
        got = memo.get(a)
        if got is not None:
            return got
        insts = instances.get(a)
        if not insts:
            memo[a] = FALSE
            return FALSE
        if a in in_progress:
            raise ValueError('cyclic ground dependency through %s' % term_to_str(a))
        in_progress.add(a)
        ors = []
        for pos, negs in insts:
            lits = [atom_lit(p) for p in pos]
            lits.extend(neg(atom_lit(n)) for n in negs)
            ors.append(nb.mk_and(lits))
        res = nb.mk_or(ors)
        in_progress.discard(a)
        memo[a] = res
        return res

    # ---- outputs
    next_lit = [atom_lit(('next', b)) for b in bases]
    legal_lit = [[atom_lit(('legal', r, m)) for m in inputs[r]] for r in roles]
    terminal_lit = atom_lit('terminal')
    goal_values, goal_lit = [], []
    for r in roles:
        vals = sorted(set(int(a[2]) for a in instances
                          if type(a) is tuple and a[0] == 'goal' and len(a) == 3 and a[1] == r))
        goal_values.append(vals)
        goal_lit.append([atom_lit(('goal', r, str(v))) for v in vals])

    # warn about legal atoms that are not declared inputs (would be unreachable through the API)
    for a in instances:
        if type(a) is tuple and a[0] == 'legal' and len(a) == 3:
            ri = role_index.get(a[1])
            if ri is None or a[2] not in input_index[ri]:
                if atom_lit(a) != FALSE:
                    raise ValueError('legal move without input declaration: ' + term_to_str(a))

    outputs = list(next_lit) + [l for ls in legal_lit for l in ls] + [terminal_lit] + \
        [l for ls in goal_lit for l in ls]

    # ---- fanout-1 flattening + DCE, then fan-in limiting, over the reachable cone
    first_gate = nb.first_gate
    gate_lits = nb.gate_lits

    def reachable(roots, lits_of):
        seen = set()
        stack = [l >> 1 for l in roots if (l >> 1) >= first_gate]
        while stack:
            n = stack.pop()
            if n in seen:
                continue
            seen.add(n)
            for l in lits_of(n):
                m = l >> 1
                if m >= first_gate and m not in seen:
                    stack.append(m)
        return seen

    live = reachable(outputs, lambda n: gate_lits[n - first_gate])
    fanout: Dict[int, int] = {}
    for n in live:
        for l in gate_lits[n - first_gate]:
            m = l >> 1
            if m >= first_gate:
                fanout[m] = fanout.get(m, 0) + 1
    for l in outputs:
        m = l >> 1
        if m >= first_gate:
            fanout[m] = fanout.get(m, 0) + 2      # outputs pin a node

    flat: Dict[int, Tuple[int, ...]] = {}

    def flatten(n: int) -> Tuple[int, ...]:
        got = flat.get(n)
        if got is not None:
            return got
        out = []
        for l in gate_lits[n - first_gate]:
            m = l >> 1
            if m >= first_gate and not (l & 1) and fanout.get(m, 0) == 1:
                out.extend(flatten(m))
            else:
                out.append(l)
        res = tuple(sorted(set(out)))
        flat[n] = res
        return res

    for n in sorted(live):
        flatten(n)
    live2 = reachable(outputs, lambda n: flat[n])

    # ---- lower: topological order, slot numbering, fan-in splitting
    order: List[int] = []
    state: Dict[int, int] = {}
    for root in sorted(live2):
        if root in state:
            continue
        stack = [(root, 0)]
        while stack:
            n, i = stack.pop()
            ls = flat[n]
            if i == 0:
                if state.get(n):
                    continue
                state[n] = 1
            while i < len(ls):
                m = ls[i] >> 1
                i += 1
                if m >= first_gate and m in live2 and state.get(m, 0) == 0:
                    stack.append((n, i))
                    stack.append((m, 0))
                    break
            else:
                state[n] = 2
                order.append(n)

    K = Propnet.MAX_FANIN
    slot_of: Dict[int, int] = {}
    out_gates: List[Tuple[int, List[int]]] = []     # (slot, literals in slot space)
    level_of_slot: Dict[int, int] = {}
    next_slot = first_gate

    def remap(l: int) -> int:
        m = l >> 1
        if m >= first_gate:
            return slot_of[m] * 2 | (l & 1)
        return l

    def lvl(l: int) -> int:
        return level_of_slot.get(l >> 1, 0)

    def emit(lits: List[int]) -> int:
        nonlocal next_slot
        s = next_slot
        next_slot += 1
        out_gates.append((s, lits))
        level_of_slot[s] = 1 + max(lvl(l) for l in lits)
        return s

    for n in order:
        lits = [remap(l) for l in flat[n]]
        # split wide gates: group by level so that early inputs are combined first
        while len(lits) > K:
            lits.sort(key=lvl)
            chunk, lits = lits[:K], lits[K:]
            lits.append(emit(chunk) * 2)
        slot_of[n] = emit(lits)

    pn = Propnet()
    pn.roles = list(roles)
    pn.bases = [term_to_str(b) for b in bases]
    pn.inputs = [[term_to_str(m) for m in inputs[r]] for r in roles]
    init = np.zeros(len(bases), dtype=np.uint8)
    for b in g['init']:
        init[base_index[b]] = 1
    pn.init_state = init
    pn.n_slots = next_slot
    pn.gate_out = np.array([s for s, _ in out_gates], dtype=np.int32)
    pn.gate_lits = [np.array(ls, dtype=np.int32) for _, ls in out_gates]
    pn.gate_level = np.array([level_of_slot[s] for s, _ in out_gates], dtype=np.int32)
    pn.next_lit = np.array([remap(l) for l in next_lit], dtype=np.int32)
    pn.legal_lit = [np.array([remap(l) for l in ls], dtype=np.int32) for ls in legal_lit]
    pn.terminal_lit = int(remap(terminal_lit))
    pn.goal_values = [np.array(v, dtype=np.int32) for v in goal_values]
    pn.goal_lit = [np.array([remap(l) for l in ls], dtype=np.int32) for ls in goal_lit]
    _make_programs(pn)
    if verbose:
        nl = sum(len(x) for x in pn.gate_lits)
        print('[propnet] bases %d inputs %s gates %d literals %d levels %d | V %d T %d TV %d' % (
            pn.n_bases, pn.role_sizes, len(pn.gate_out), nl,
            int(pn.gate_level.max()) if len(pn.gate_level) else 0,
            len(pn.prog['V']), len(pn.prog['T']), len(pn.prog['TV'])), file=sys.stderr)
    return pn


def _cone(pn: Propnet, roots) -> np.ndarray:
    first_gate = 1 + pn.n_bases + sum(pn.role_sizes)
    gate_index = {int(s): i for i, s in enumerate(pn.gate_out)}
    seen = set()
    stack = [int(l) >> 1 for l in roots if (int(l) >> 1) >= first_gate]
    while stack:
        s = stack.pop()
        if s in seen:
            continue
        seen.add(s)
        for l in pn.gate_lits[gate_index[s]]:
            m = int(l) >> 1
            if m >= first_gate and m not in seen:
                stack.append(m)
    idx = sorted((gate_index[s] for s in seen), key=lambda i: (int(pn.gate_level[i]), i))
    return np.array(idx, dtype=np.int32)


def _make_programs(pn: Propnet):
    v_roots = [l for ls in pn.legal_lit for l in ls] + [pn.terminal_lit] + \
        [l for ls in pn.goal_lit for l in ls]
    lt_roots = [l for ls in pn.legal_lit for l in ls] + [pn.terminal_lit]
    g_roots = [l for ls in pn.goal_lit for l in ls]
    V = _cone(pn, v_roots)
    T = _cone(pn, pn.next_lit)
    inV = set(V.tolist())
    TV = np.array([i for i in T.tolist() if i not in inV], dtype=np.int32)
    LT = _cone(pn, lt_roots)
    inLT = set(LT.tolist())
    G = _cone(pn, g_roots)
    GL = np.array([i for i in G.tolist() if i not in inLT], dtype=np.int32)
    TL = np.array([i for i in T.tolist() if i not in inLT], dtype=np.int32)
    pn.prog = {'V': V, 'T': T, 'TV': TV, 'LT': LT, 'GL': GL, 'TL': TL}


# --------------------------------------------------------------------------------------------
def compile_kif_text(text: str, verbose: bool = False) -> Propnet:
    from .ground import Grounder
    from .parser import parse_rules
    g = Grounder(parse_rules(text), verbose=verbose).ground()
    return build_propnet(g, verbose=verbose)


def compile_kif(path: str, verbose: bool = False) -> Propnet:
    with open(path) as f:
        return compile_kif_text(f.read(), verbose=verbose)


# --------------------------------------------------------------------------------------------
def save_propnet(pn: Propnet, path: str):
    import json
    meta = {
        'format': 'gzero-b200-propnet-1',
        'roles': pn.roles, 'bases': pn.bases, 'inputs': pn.inputs,
        'n_slots': pn.n_slots, 'terminal_lit': pn.terminal_lit,
    }
    offs = np.zeros(len(pn.gate_lits) + 1, dtype=np.int64)
    for i, ls in enumerate(pn.gate_lits):
        offs[i + 1] = offs[i] + len(ls)
    arrays = {
        'meta': np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
        'init_state': pn.init_state,
        'gate_out': pn.gate_out, 'gate_level': pn.gate_level,
        'gate_lit_offsets': offs,
        'gate_lit_data': (np.concatenate(pn.gate_lits) if pn.gate_lits
                          else np.zeros(0, np.int32)).astype(np.int32),
        'next_lit': pn.next_lit,
    }
    for r in range(len(pn.roles)):
        arrays['legal_lit_%d' % r] = pn.legal_lit[r]
        arrays['goal_values_%d' % r] = pn.goal_values[r]
        arrays['goal_lit_%d' % r] = pn.goal_lit[r]
    for k, v in pn.prog.items():
        arrays['prog_' + k] = v
    np.savez_compressed(path, **arrays)


def load_propnet(path: str) -> Propnet:
    import json
    z = np.load(path)
    meta = json.loads(bytes(z['meta']).decode())
    pn = Propnet()
    pn.roles = meta['roles']
    pn.bases = meta['bases']
    pn.inputs = meta['inputs']
    pn.n_slots = meta['n_slots']
    pn.terminal_lit = meta['terminal_lit']
    pn.init_state = z['init_state']
    pn.gate_out = z['gate_out']
    pn.gate_level = z['gate_level']
    offs = z['gate_lit_offsets']
    data = z['gate_lit_data']
    pn.gate_lits = [data[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
    pn.next_lit = z['next_lit']
    nr = len(pn.roles)
    pn.legal_lit = [z['legal_lit_%d' % r] for r in range(nr)]
    pn.goal_values = [z['goal_values_%d' % r] for r in range(nr)]
    pn.goal_lit = [z['goal_lit_%d' % r] for r in range(nr)]
    pn.prog = {k[5:]: z[k] for k in z.files if k.startswith('prog_')}
    if 'TL' not in pn.prog:
        in_lt = set(pn.prog['LT'].tolist())
        pn.prog['TL'] = np.array([i for i in pn.prog['T'].tolist() if i not in in_lt], dtype=np.int32)
    return pn
