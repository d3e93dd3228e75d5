#!/usr/bin/env python
"""Golden vectors from the REFERENCE'S OWN Python transformer, generated in the build container:

    python tests/golden/make_reference_python_golden.py        # needs /root/reference

`ggpzero.nn.bases` (src/ggpzero/nn/bases.py) and `ggpzero.defs.{datadesc,gamedesc}` are imported unmodified from
/root/reference/src. What they import from packages that are not in this container is stubbed with the surface
they touch: `ggplib.util.symbols.SymbolFactory.symbolize` (an s-expression reader returning nested `ListTerm`s — a
tuple subclass — of `Term`s — a str subclass), `ggplib.db.lookup.GameInfo` (a type to isinstance against), `ggplib.util.log`, and
`ggpzero_interface.buf_to_tuple_reverse_bytes` (np.unpackbits). The game model (roles, bases, actions as GDL strings)
comes from this build's compiler — the base / action ORDER is this build's (unpinned by the reference, DESIGN.md
§2); everything computed FROM that order is the reference's code:

  * `GdlBasesTransformer.__init__` / `init_spaces` / `create_board_space` / `create_control_space`
    (bases.py:62-240) with the reference's `gamedesc.Games()` descriptions -> board / control tables,
  * `state_to_channels(state, prev_states)` (bases.py:242-287) on positions reached by random play,
  * `policy_to_array`, `value_to_array` (plain and the draw head's, bases.py:305-335) and `check_sample`
    (bases.py:289-303) on sample records,

  * `ggpzero.util.symmetry` (src/ggpzero/util/symmetry.py, imported unmodified as well): `create_translator` with the
    reference's `gamedesc.GameSymmetries`, `Prescription`, and for every prescription entry the destination of every
    base (`translate_basestate_helper`; skipped bases stay) and of every action (`translate_action`), plus whole
    states through `translate_basestate` -> tests/golden/refpy_symmetry_<game>.json.gz,

stored in tests/golden/refpy_<game>.json.gz and checked by tests/test_reference_python_golden.py against
galvanise_zero_b200/bases.py, the C oracle encoder (oracle/encoder_ref.c), oracle/training_ref.py and — on a GPU —
gz_encode / gz_records_to_training."""
import gzip
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

GAMES = ['reversi', 'breakthrough', 'breakthroughSmall', 'connect6', 'amazons_10x10', 'chess_15d', 'hexLG13']


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class Term(str):
        pass

    class ListTerm(tuple):
        pass

    class SymbolFactory(object):
        def to_symbols(self, s):
            toks = s.replace('(', ' ( ').replace(')', ' ) ').split()

            def parse(i):
                out = []
                while i < len(toks):
                    t = toks[i]
                    if t == '(':
                        sub, i = parse(i + 1)
                        out.append(ListTerm(sub))
                    elif t == ')':
                        return out, i + 1
                    else:
                        out.append(Term(t))
                        i += 1
                return out, i
            return parse(0)[0]

        def symbolize(self, s):
            return self.to_symbols(s)[0]

    class GameInfo(object):
        pass

    class _Log(object):
        def __getattr__(self, k):
            return lambda *a, **kw: None

    mod('ggplib').__path__ = []
    mod('ggplib.util', log=_Log()).__path__ = []
    mod('ggplib.util.symbols', SymbolFactory=SymbolFactory, Term=Term, ListTerm=ListTerm)
    mod('ggplib.util.log')
    mod('ggplib.db').__path__ = []
    mod('ggplib.db.lookup', GameInfo=GameInfo)
    mod('ggpzero_interface',
        buf_to_tuple_reverse_bytes=lambda buf: tuple(bool(x) for x in np.unpackbits(np.frombuffer(buf, dtype=np.uint8))))
    return GameInfo


def main(games):
    GameInfo = install_stubs()
    sys.path.insert(0, '/root/reference/src')
    from ggpzero.defs import datadesc as ref_datadesc
    from ggpzero.nn import bases as ref_bases
    from galvanise_zero_b200 import lookup, tables
    from galvanise_zero_b200.defs import datadesc as my_datadesc
    from oracle.pnref import RefStateMachine
    import test_training_batch as ttb              # synthetic_records (records over reachable states)
    from oracle import training_ref as tref

    for game in games:
        mine = lookup.by_name(game)
        gi = GameInfo()
        gi.game = game
        gi.model = types.SimpleNamespace(roles=list(mine.model.roles), bases=list(mine.model.bases),
                                         actions=[list(a) for a in mine.model.actions])
        gd = my_datadesc.generation_for(game)
        ref_gd = ref_datadesc.GenerationDescription(game=game, name='b200_synthetic', channel_last=False,
                                                    multiple_policy_heads=True, num_previous_states=gd.num_previous_states,
                                                    draw_head=gd.draw_head)
        cls = ref_bases.GdlBasesTransformer_Draws if gd.draw_head else ref_bases.GdlBasesTransformer
        tr = cls(gi, ref_gd)
        out = dict(game=game, n_bases=len(gi.model.bases), policy_dist_count=list(tr.policy_dist_count),
                   num_channels=tr.num_channels, num_rows=tr.num_rows, num_cols=tr.num_cols,
                   raw_channels_per_state=tr.raw_channels_per_state,
                   num_of_controls_channels=tr.num_of_controls_channels, num_rewards=tr.num_rewards,
                   board_space=[[b.base_indx, b.channel_id, b.x_idx, b.y_idx] for b in tr.board_space],
                   control_space=[[c.base_indx, c.channel_id, float(c.value)] for c in tr.control_space])
        # ---- sample records over reachable positions, then the reference's per-sample functions
        sm = RefStateMachine(tables.blob_by_name(game))
        my_tr = types.SimpleNamespace(policy_dist_count=list(tr.policy_dist_count))
        n = 6 if game in ('hexLG13', 'chess_15d') else 12
        recs = ttb.synthetic_records(my_tr, sm, n, seed=4242, corrupt={1})
        nb = len(gi.model.bases)
        rows = []
        for r in recs:
            state, prevs, policies, final_score = tref.decode_record(r, sm.state_words, nb, len(tr.policy_dist_count))
            state = [int(x) for x in state]
            prevs = [[int(x) for x in p] for p in prevs]
            planes = tr.state_to_channels(state, prevs)
            assert planes.shape == (tr.num_channels, tr.num_cols, tr.num_rows), planes.shape
            flat = np.asarray(planes, dtype=np.float32).reshape(-1)
            nz = np.nonzero(flat)[0]
            fs = [max(0.0, min(1.0, v)) for v in final_score]                     # datacache.py:716-718
            pol = [tr.policy_to_array(p, ri) for ri, p in enumerate(policies)]
            try:
                val = [float(x) for x in tr.value_to_array(fs)]
            except AssertionError:
                val = None
            sample = ref_datadesc.Sample(state=state, prev_states=prevs, policies=policies, final_score=fs)
            try:
                tr.check_sample(sample)
                ok = True
            except AssertionError:
                ok = False
            rows.append(dict(record=[int(x) for x in r],
                             planes_nonzero=[[int(i), float(flat[i])] for i in nz],
                             policy_nonzero=[[[int(i), float(a[i])] for i in np.nonzero(a)[0]] for a in pol],
                             value=val, check_sample=ok))
        # ---- the reference's symmetry translator
        from ggpzero.util import symmetry as ref_symmetry
        from ggpzero.defs import gamedesc as ref_gamedesc
        gds = ref_gamedesc.GameSymmetries()
        if hasattr(gds, game):
            desc = getattr(gds, game)()
            t = ref_symmetry.create_translator(gi, tr.game_desc, desc)
            pres = [[bool(r), int(k)] for r, k in ref_symmetry.Prescription(desc)]
            base_dest, action_dest = [], []
            for refl, rot in pres:
                dest = []
                for idx, terms in enumerate(t.base_symbols):
                    if terms[0] in t.skip_base_root_terms:
                        dest.append(idx)
                    else:
                        dest.append(int(t.translate_basestate_helper(terms, refl, rot)))
                base_dest.append(dest)
                action_dest.append([[int(t.translate_action(ri, a, refl, rot)) for a in range(len(gi.model.actions[ri]))]
                                    for ri in range(len(gi.model.roles))])
            states = []
            for r in recs[:3]:
                bits = [int(x) for x in tref.decode_record(r, sm.state_words, nb, len(tr.policy_dist_count))[0]]
                states.append(dict(bits_set=[i for i, b in enumerate(bits) if b],
                                   translated=[[i for i, b in enumerate(t.translate_basestate(bits, refl, rot)) if b]
                                               for refl, rot in pres]))
            sym = dict(game=game, prescription=pres, base_dest=base_dest, action_dest=action_dest, states=states)
            spath = os.path.join(HERE, 'refpy_symmetry_%s.json.gz' % game)
            with gzip.open(spath, 'wt') as f:
                json.dump(sym, f, separators=(',', ':'))
            print(game, 'symmetry', pres, os.path.getsize(spath), 'bytes')
        out['state_words'] = sm.state_words
        out['samples'] = rows
        path = os.path.join(HERE, 'refpy_%s.json.gz' % game)
        with gzip.open(path, 'wt') as f:
            json.dump(out, f, separators=(',', ':'))
        print(game, 'board', len(out['board_space']), 'control', len(out['control_space']), 'samples', len(rows),
              os.path.getsize(path), 'bytes')


if __name__ == '__main__' and not any(a in sys.argv for a in ('--configs', '--gendata', '--variants')):
    main(sys.argv[1:] or GAMES)


def configs_golden():
    """attr.asdict of the reference's own config classes and templates (src/ggpzero/defs/{confs,templates,datadesc}.py,
    imported unmodified) -> tests/golden/refpy_configs.json: the config surface the engine must accept verbatim
    (SURVEY.md appendix A)."""
    import attr
    install_stubs()
    sys.path.insert(0, '/root/reference/src')
    from ggpzero.defs import confs, datadesc, templates
    out = dict(
        PUCTEvaluatorConfig=attr.asdict(confs.PUCTEvaluatorConfig()),
        PUCTPlayerConfig=attr.asdict(confs.PUCTPlayerConfig()),
        SelfPlayConfig=attr.asdict(confs.SelfPlayConfig()),
        base_puct_config=attr.asdict(templates.base_puct_config()),
        base_puct_config_overridden=attr.asdict(templates.base_puct_config(dirichlet_noise_pct=0.15, random_scale=0.75)),
        selfplay_config_template=attr.asdict(templates.selfplay_config_template()),
        fields=dict(Sample=[a.name for a in attr.fields(datadesc.Sample)],
                    GenerationDescription=[a.name for a in attr.fields(datadesc.GenerationDescription)],
                    GenerationSamples=[a.name for a in attr.fields(datadesc.GenerationSamples)]),
        GenerationDescription=attr.asdict(datadesc.GenerationDescription()),
        default_generation_desc=attr.asdict(templates.default_generation_desc('reversi', multiple_policy_heads=True,
                                                                               num_previous_states=1)))
    out['default_generation_desc'].pop('date_created')          # (a timestamp)
    from ggpzero.defs import gamedesc
    out['gamedesc'] = {g: attr.asdict(getattr(gamedesc.Games(), g)()) for g in GAMES}
    out['game_symmetries'] = {g: attr.asdict(getattr(gamedesc.GameSymmetries(), g)()) for g in GAMES
                              if hasattr(gamedesc.GameSymmetries(), g)}
    import types as _types
    shape = _types.SimpleNamespace(role_count=2, num_rows=8, num_cols=8, num_channels=5, policy_dist_count=[65, 65])
    out['nn_model_config_template'] = {h: attr.asdict(templates.nn_model_config_template('reversi', h, shape))
                                       for h in ('small', 'medium', 'large')}
    out['nn_model_config_template']['NNModelConfig'] = attr.asdict(confs.NNModelConfig())
    path = os.path.join(HERE, 'refpy_configs.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=1, sort_keys=False)
    print('configs', os.path.getsize(path), 'bytes')


if __name__ == '__main__' and '--configs' in sys.argv:
    configs_golden()


def gendata_golden():
    """A generation of samples serialised by the reference's OWN writer (`attrutil.attr_to_json(GenerationSamples)`,
    src/ggpzero/util/attrutil.py, what `ServerBroker.save_sample_data` writes, distributed/server.py:300-322) ->
    tests/golden/refpy_gendata_reversi.json, with the records the samples came from."""
    import attr
    install_stubs()
    sys.path.insert(0, '/root/reference/src')
    from ggpzero.defs import datadesc as ref_datadesc
    from ggpzero.util import attrutil
    from galvanise_zero_b200 import bases, gendata, lookup, tables
    from galvanise_zero_b200.defs import datadesc
    from oracle.pnref import RefStateMachine
    import test_training_batch as ttb
    game = 'reversi'
    tr = bases.GdlBasesTransformer(lookup.by_name(game), datadesc.generation_for(game))
    sm = RefStateMachine(tables.blob_by_name(game))
    recs = ttb.synthetic_records(tr, sm, 5, seed=99)
    f = recs.view(np.float32)
    for i in range(len(recs)):             # values that survive the file's 5-decimal floats
        nc = recs[i, 11]
        p0 = 24 + 2 * sm.state_words
        f[i, p0 + 1:p0 + 2 * nc:2] = np.round(f[i, p0 + 1:p0 + 2 * nc:2].astype(np.float64), 5).astype(np.float32)
        f[i, 12:14] = np.round(f[i, 12:14].astype(np.float64), 5).astype(np.float32)
    mine = [gendata.encode_sample_states(datadesc.Sample(**d)) for d in ttb._sample_dicts(recs, tr, sm)]
    for k, s in enumerate(mine):
        s.match_identifier = 'golden_%d' % k
        s.resultant_puct_score = [0.25, 0.75]
        s.policies = [[(int(m), float('%.5f' % p)) for m, p in pol] for pol in s.policies]
        s.final_score = [float('%.5f' % v) for v in s.final_score]
    ref_samples = [ref_datadesc.Sample(**attr.asdict(s, recurse=False)) for s in mine]
    ref_gen = ref_datadesc.GenerationSamples(game=game, date_created='2026/01/01 00:00', with_generation='golden_gen_7',
                                             num_samples=len(ref_samples), samples=ref_samples)
    ref_json = attrutil.attr_to_json(ref_gen, pretty=False)
    my_gen = datadesc.GenerationSamples(game=game, date_created='2026/01/01 00:00', with_generation='golden_gen_7',
                                        num_samples=len(mine), samples=mine)
    my_json = gendata.generation_to_json(my_gen)
    # (the reference's reader, attrutil.json_to_attr, pops keys from the dict it iterates over — Python 2 only — so
    # it cannot be run here; the two files are compared as JSON documents instead)
    assert json.loads(my_json) == json.loads(ref_json), 'the two writers disagree'
    path = os.path.join(HERE, 'refpy_gendata_reversi.json')
    with open(path, 'w') as fo:
        json.dump(dict(reference_json=ref_json, records=[[int(x) for x in r] for r in recs]), fo)
    print('gendata', os.path.getsize(path), 'bytes; this build writer produces the same JSON document')


if __name__ == '__main__' and '--gendata' in sys.argv:
    gendata_golden()


def variants_golden():
    """The reference's `state_to_channels` for other generation shapes — no previous state, two previous states,
    channels last — on reversi and connect6 -> tests/golden/refpy_variants.json.gz."""
    GameInfo = install_stubs()
    sys.path.insert(0, '/root/reference/src')
    from ggpzero.defs import datadesc as ref_datadesc
    from ggpzero.nn import bases as ref_bases
    from galvanise_zero_b200 import lookup, tables
    from oracle.pnref import RefStateMachine
    out = {}
    for game in ('reversi', 'connect6'):
        mine = lookup.by_name(game)
        gi = GameInfo()
        gi.game = game
        gi.model = types.SimpleNamespace(roles=list(mine.model.roles), bases=list(mine.model.bases),
                                         actions=[list(a) for a in mine.model.actions])
        sm = RefStateMachine(tables.blob_by_name(game))
        n = 5
        s = np.zeros((n, sm.state_words), dtype=np.uint32)
        l = np.zeros((n, sm.legal_words), dtype=np.uint32)
        t = np.ones(n, dtype=np.uint8)
        g = np.zeros((n, sm.n_roles), dtype=np.uint8)
        hist = []
        for step in range(8):
            sm.playout_step(0, 2024, step, s, l, t, g, None)
            hist.append(s.copy())
        nb = len(gi.model.bases)
        bits = lambda w: [int(x) for x in tables.unpack_state_bits(w, nb)]      # noqa: E731
        cases = []
        for num_prev, channel_last in ((0, False), (2, False), (1, True), (3, False)):
            gd = ref_datadesc.GenerationDescription(game=game, name='v', channel_last=channel_last,
                                                    multiple_policy_heads=True, num_previous_states=num_prev)
            tr = ref_bases.GdlBasesTransformer(gi, gd)
            rows = []
            for i in range(n):
                for have in sorted({0, min(1, num_prev), num_prev}):
                    prevs = [bits(hist[6 - k][i]) for k in range(have)]
                    planes = np.asarray(tr.state_to_channels(bits(hist[7][i]), prevs), dtype=np.float32)
                    flat = planes.reshape(-1)
                    rows.append(dict(i=i, have=have, shape=list(planes.shape),
                                     nonzero=[[int(j), float(flat[j])] for j in np.nonzero(flat)[0]]))
            cases.append(dict(num_previous_states=num_prev, channel_last=channel_last, num_channels=tr.num_channels, rows=rows))
        out[game] = dict(states=[[[int(x) for x in hist[k][i]] for i in range(n)] for k in range(8)], cases=cases)
    path = os.path.join(HERE, 'refpy_variants.json.gz')
    with gzip.open(path, 'wt') as f:
        json.dump(out, f, separators=(',', ':'))
    print('variants', os.path.getsize(path), 'bytes')


if __name__ == '__main__' and '--variants' in sys.argv:
    variants_golden()
