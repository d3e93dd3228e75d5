"""Golden vectors computed by the reference's OWN Python transformer (src/ggpzero/nn/bases.py imported unmodified
from /root/reference by tests/golden/make_reference_python_golden.py, ggplib stubbed): board / control tables of
GdlBasesTransformer, state_to_channels planes, policy_to_array / value_to_array rows and check_sample verdicts for
sample records over reachable positions of all seven games.

CPU: galvanise_zero_b200/bases.py (tables, planes, rows, verdicts), the C oracle encoder and oracle/training_ref.py
reproduce them exactly — which pins those restatements to outputs of the reference's code rather than to a reading
of it. GPU: gz_records_to_training (encoder + rec_unpack_kernel) reproduces them through the C-ABI."""
import gzip
import json
import os

import numpy as np
import pytest

from galvanise_zero_b200 import bases, cppinterface, lookup
from galvanise_zero_b200.defs import datadesc
from oracle import training_ref as tref
from oracle.pnref import RefEncoder

HERE = os.path.dirname(os.path.abspath(__file__))
GAMES = ['reversi', 'breakthrough', 'breakthroughSmall', 'connect6', 'amazons_10x10', 'chess_15d', 'hexLG13']


def _golden(game):
    with gzip.open(os.path.join(HERE, 'golden', 'refpy_%s.json.gz' % game), 'rt') as f:
        return json.load(f)


def _dense(pairs, size):
    a = np.zeros(size, dtype=np.float32)
    for i, v in pairs:
        a[i] = np.float32(v)
    return a


def _setup(game):
    tr = bases.GdlBasesTransformer(lookup.by_name(game), datadesc.generation_for(game))
    return tr, cppinterface.create_c_transformer(tr)


@pytest.mark.parametrize('game', GAMES)
def test_transformer_tables_are_the_references(game):
    g = _golden(game)
    tr, ct = _setup(game)
    assert tr.num_bases == g['n_bases'] and list(tr.policy_dist_count) == g['policy_dist_count']
    assert (tr.num_channels, tr.num_rows, tr.num_cols) == (g['num_channels'], g['num_rows'], g['num_cols'])
    assert tr.raw_channels_per_state == g['raw_channels_per_state']
    assert tr.num_of_controls_channels == g['num_of_controls_channels'] and tr.num_rewards == g['num_rewards']
    assert [[b.base_indx, b.channel_id, b.x_idx, b.y_idx] for b in tr.board_space] == g['board_space']
    assert [[c.base_indx, c.channel_id, float(c.value)] for c in tr.control_space] == g['control_space']


@pytest.mark.parametrize('game', GAMES)
def test_planes_rows_and_verdicts_are_the_references(game):
    g = _golden(game)
    tr, ct = _setup(game)
    sw, nb, widths, nv = g['state_words'], g['n_bases'], g['policy_dist_count'], g['num_rewards']
    recs = np.array([s['record'] for s in g['samples']], dtype=np.int32)
    total = tr.num_channels * tr.channel_size
    want_planes = np.stack([_dense(s['planes_nonzero'], total) for s in g['samples']])
    want_pol = [np.stack([_dense(s['policy_nonzero'][ri], widths[ri]) for s in g['samples']]) for ri in range(len(widths))]
    # the oracle restatement (C encoder + training_ref) straight from the records
    planes, pols, vals, bad = tref.records_to_training(recs, sw, nb, widths, nv, RefEncoder.from_c_transformer(ct))
    assert np.array_equal(planes, want_planes)
    assert all(np.array_equal(a, b) for a, b in zip(pols, want_pol))
    assert bad == sum(1 for s in g['samples'] if not s['check_sample']) and bad >= 1
    for i, s in enumerate(g['samples']):
        if s['value'] is not None:
            assert vals[i].tolist() == [np.float32(v) for v in s['value']]
    # the Python mirror of the transformer on the decoded samples
    for i, s in enumerate(g['samples']):
        state, prevs, policies, final_score = tref.decode_record(recs[i], sw, nb, len(widths))
        assert np.array_equal(tr.state_to_channels(state, prevs).reshape(-1), want_planes[i])
        fs = [max(0.0, min(1.0, v)) for v in final_score]
        for ri in range(len(widths)):
            assert np.array_equal(tr.policy_to_array(policies[ri], ri), want_pol[ri][i])
        if s['value'] is not None:
            assert tr.value_to_array(fs).tolist() == [np.float32(v) for v in s['value']]
        sample = datadesc.Sample(state=[int(x) for x in state], prev_states=[[int(x) for x in p] for p in prevs],
                                 policies=policies, final_score=fs)
        if s['check_sample']:
            tr.check_sample(sample)
        else:
            with pytest.raises(AssertionError):
                tr.check_sample(sample)


@pytest.mark.gpu
@pytest.mark.parametrize('game', GAMES)
def test_device_arrays_are_the_references(game):
    import torch
    from galvanise_zero_b200.datacache import TrainingBatcher
    g = _golden(game)
    tr, ct = _setup(game)
    widths = g['policy_dist_count']
    recs = np.array([s['record'] for s in g['samples']], dtype=np.int32)
    total = tr.num_channels * tr.channel_size
    want_planes = np.stack([_dense(s['planes_nonzero'], total) for s in g['samples']])
    out = TrainingBatcher(tr, ct, g['state_words'], device=0).arrays(torch.from_numpy(recs).cuda(), check=True)
    assert np.array_equal(out['planes'].reshape(len(recs), -1).cpu().numpy(), want_planes)
    for ri in range(len(widths)):
        want = np.stack([_dense(s['policy_nonzero'][ri], widths[ri]) for s in g['samples']])
        assert np.array_equal(out['policies'][ri].cpu().numpy(), want)
    got_v = out['value'].cpu().numpy()
    for i, s in enumerate(g['samples']):
        if s['value'] is not None:
            assert got_v[i].tolist() == [np.float32(v) for v in s['value']]
    assert out['bad'] == sum(1 for s in g['samples'] if not s['check_sample'])


SYM_GAMES = ['reversi', 'breakthrough', 'breakthroughSmall', 'connect6', 'amazons_10x10', 'hexLG13']


@pytest.mark.parametrize('game', SYM_GAMES)
def test_symmetry_tables_are_the_references(game):
    """Prescription, the destination of every base and of every action under every prescription entry, and whole
    translated states: computed by the reference's own `ggpzero.util.symmetry` (Translator / create_translator /
    Prescription imported unmodified) with the reference's `gamedesc.GameSymmetries`."""
    from galvanise_zero_b200 import symmetry
    from galvanise_zero_b200.defs import gamedesc
    with gzip.open(os.path.join(HERE, 'golden', 'refpy_symmetry_%s.json.gz' % game), 'rt') as f:
        g = json.load(f)
    gi = lookup.by_name(game)
    desc = gamedesc.get_symmetries_desc(game)
    t = symmetry.create_translator(gi, getattr(gamedesc.Games(), game)(), desc)
    pres = [[bool(r), int(k)] for r, k in symmetry.Prescription(desc)]
    assert pres == g['prescription']
    nb = len(gi.model.bases)
    for k, (refl, rot) in enumerate(pres):
        assert np.asarray(t.base_permutation(refl, rot)).tolist() == g['base_dest'][k]
        for ri in range(len(gi.model.roles)):
            assert np.asarray(t.action_permutation(ri, refl, rot)).tolist() == g['action_dest'][k][ri]
        assert t.translate_action(0, 1, refl, rot) == g['action_dest'][k][0][1]
        assert t.translate_base_index(nb - 1, refl, rot) == g['base_dest'][k][nb - 1]
        for st in g['states']:
            bits = [0] * nb
            for i in st['bits_set']:
                bits[i] = 1
            got = t.translate_basestate(bits, refl, rot)
            assert [i for i, b in enumerate(got) if b] == st['translated'][k]
    assert gamedesc.get_symmetries_desc('chess_15d') is None        # (the reference has none for chess either)


def test_config_surface_is_the_references():
    """attr.asdict of the reference's own PUCTEvaluatorConfig / PUCTPlayerConfig / SelfPlayConfig, of its templates
    (base_puct_config, selfplay_config_template) and the field lists of its datadesc classes — keys, order and
    defaults — equal this build's: the dicts `start_self_play` / `Player` receive are the reference's verbatim."""
    import attr
    from galvanise_zero_b200.defs import confs
    with open(os.path.join(HERE, 'golden', 'refpy_configs.json')) as f:
        g = json.load(f)

    def same(ref, mine):
        assert list(ref) == list(mine) and ref == mine

    same(g['PUCTEvaluatorConfig'], attr.asdict(confs.PUCTEvaluatorConfig()))
    same(g['PUCTPlayerConfig'], attr.asdict(confs.PUCTPlayerConfig()))
    same(g['SelfPlayConfig'], attr.asdict(confs.SelfPlayConfig()))
    same(g['base_puct_config'], attr.asdict(confs.base_puct_config()))
    same(g['base_puct_config_overridden'], attr.asdict(confs.base_puct_config(dirichlet_noise_pct=0.15, random_scale=0.75)))
    same(g['selfplay_config_template'], attr.asdict(confs.selfplay_config_template()))
    same(g['GenerationDescription'], attr.asdict(datadesc.GenerationDescription()))
    for name, fields in g['fields'].items():
        assert [a.name for a in attr.fields(getattr(datadesc, name))] == fields
    gd = attr.asdict(datadesc.generation_for('reversi', draw_head=False))
    for k, v in g['default_generation_desc'].items():
        if k not in ('name',):
            assert gd[k] == v, k
    # the game descriptions and symmetry descriptions of defs/gamedesc.py for the compiled games
    from galvanise_zero_b200.defs import gamedesc
    for game, ref in g['gamedesc'].items():
        same(ref, json.loads(json.dumps(attr.asdict(getattr(gamedesc.Games(), game)()))))
    for game, ref in g['game_symmetries'].items():
        same(ref, json.loads(json.dumps(attr.asdict(getattr(gamedesc.GameSymmetries(), game)()))))
    assert sorted(g['game_symmetries']) == sorted(x for x in g['gamedesc'] if gamedesc.get_symmetries_desc(x) is not None)
    # the network shapes of templates.nn_model_config_template (the benchmarked "small" tower: 64 filters x 5 blocks)
    import types
    from galvanise_zero_b200 import nn as gznn
    shape = types.SimpleNamespace(role_count=2, num_rows=8, num_cols=8, num_channels=5, policy_dist_count=[65, 65])
    for hint in ('small', 'medium', 'large'):
        same(g['nn_model_config_template'][hint], attr.asdict(gznn.nn_model_config_template(shape, hint)))
    same(g['nn_model_config_template']['NNModelConfig'], attr.asdict(gznn.NNModelConfig()))


def test_other_generation_shapes_are_the_references():
    """No previous state, two / three previous states (each in its own channel group, fewer handed in than the
    generation has), channels last: `bases.GdlBasesTransformer.state_to_channels` equals the reference's own; with
    at most one previous state the C oracle encoder (the engine's toChannels restatement) agrees as well."""
    from galvanise_zero_b200 import tables
    with gzip.open(os.path.join(HERE, 'golden', 'refpy_variants.json.gz'), 'rt') as f:
        g = json.load(f)
    for game, d in g.items():
        gi = lookup.by_name(game)
        nb = len(gi.model.bases)
        hist = np.array(d['states'], dtype=np.uint32)                       # [ply][game][word]
        bits = lambda w: [int(x) for x in tables.unpack_state_bits(w, nb)]  # noqa: E731
        for case in d['cases']:
            gd = datadesc.GenerationDescription(game=game, name='v', channel_last=case['channel_last'],
                                                multiple_policy_heads=True, num_previous_states=case['num_previous_states'])
            tr = bases.GdlBasesTransformer(gi, gd)
            assert tr.num_channels == case['num_channels']
            enc = RefEncoder.from_c_transformer(cppinterface.create_c_transformer(tr))
            for row in case['rows']:
                i, have = row['i'], row['have']
                prevs = [bits(hist[6 - k][i]) for k in range(have)]
                planes = tr.state_to_channels(bits(hist[7][i]), prevs)
                assert list(planes.shape) == row['shape']
                want = _dense(row['nonzero'], planes.size)
                assert np.array_equal(planes.reshape(-1), want)
                if not case['channel_last'] and case['num_previous_states'] <= 1:
                    pw = hist[6][i][None, None, :] if have else None
                    got = enc.encode(hist[7][i][None, :], pw)[0]
                    assert np.array_equal(got, want)
