"""Sample records -> training arrays (the trainer side of the sample format):
DataCache.sync's per-sample work and DataCache.augment_data (src/ggpzero/nn/datacache.py:597-760) with
GdlBasesTransformer.check_sample / policy_to_array / value_to_array (src/ggpzero/nn/bases.py:289-335).

CPU: the oracle restatement (oracle/training_ref.py, straight from records) agrees with the Python transformer
mirror (galvanise_zero_b200/bases.py) fed with decoded sample dicts; the host value / policy / check functions on
the reference's edge cases (draw head, clamping, bad probabilities).
GPU: gz_records_to_training through galvanise_zero_b200.datacache.TrainingBatcher is bit-exact with the oracle
(float32 planes; bf16 = rounded float32), with and without augmentation, on synthetic records that exercise the
edge cases and on records a self-play run produced."""
import random
import types

import numpy as np
import pytest

from galvanise_zero_b200 import bases, cppinterface, lookup, symmetry, tables
from galvanise_zero_b200.defs import datadesc, gamedesc
from oracle import training_ref as tref
from oracle.pnref import RefEncoder, RefStateMachine

H = tref.SR_HEADER


def _setup(game):
    gi = lookup.by_name(game)
    tr = bases.GdlBasesTransformer(gi, datadesc.generation_for(game))
    ct = cppinterface.create_c_transformer(tr)
    sm = RefStateMachine(tables.blob_by_name(game))
    return gi, tr, ct, sm


def _states(sm, n, plies, seed):
    """n reachable states and their predecessors (random play), packed rows."""
    s = np.zeros((n, sm.state_words), dtype=np.uint32)
    l = np.zeros((n, sm.legal_words), dtype=np.uint32)
    t = np.ones(n, dtype=np.uint8)
    g = np.zeros((n, sm.n_roles), dtype=np.uint8)
    prev = s.copy()
    for step in range(plies):
        prev = s.copy()
        sm.playout_step(0, seed, step, s, l, t, g, None)
    return s, prev


def synthetic_records(tr, sm, n, seed, corrupt=()):
    """Records with the engine's layout over real states: a lead role with 1..min(width, 40) distinct actions and
    normalised float32 probabilities, the other role's single move, final scores incl. draws and out-of-range
    values (the trainer clamps them), some roots without a previous state. `corrupt`: record ids that get a
    probability vector summing to 1.5 (check_sample must flag exactly those)."""
    rng = np.random.RandomState(seed)
    sw = sm.state_words
    widths = tr.policy_dist_count
    rw = H + 2 * sw + 2 * max(widths)
    cur, prev = _states(sm, n, 7, seed)
    recs = np.zeros((n, rw), dtype=np.int32)
    f = recs.view(np.float32)
    scores = [(1.0, 0.0), (0.0, 1.0), (0.5, 0.5), (0.495, 0.505), (1.02, -0.02), (0.75, 0.25)]
    for i in range(n):
        lead = int(rng.randint(0, len(widths)))
        nc = int(rng.randint(1, min(widths[lead], 40) + 1))
        moves = rng.choice(widths[lead], size=nc, replace=False)
        p = rng.rand(nc).astype(np.float64) + 0.01
        p = (p / p.sum()).astype(np.float32)
        if i in corrupt:
            p = (p * np.float32(1.5)).astype(np.float32)
        recs[i, tref.SR_DEPTH] = i % 50
        recs[i, tref.SR_LEAD] = lead
        recs[i, tref.SR_N_PREV] = 0 if i % 5 == 0 else 1
        recs[i, tref.SR_N_CHILDREN] = nc
        fs = scores[i % len(scores)]
        for r in range(len(widths)):
            f[i, tref.SR_FINAL_SCORE + r] = fs[r]
            recs[i, tref.SR_OTHER_MOVES + r] = int(rng.randint(0, widths[r]))
        recs[i, H:H + sw] = cur[i].view(np.int32)
        if recs[i, tref.SR_N_PREV]:
            recs[i, H + sw:H + 2 * sw] = prev[i].view(np.int32)
        recs[i, H + 2 * sw:H + 2 * sw + 2 * nc:2] = moves
        f[i, H + 2 * sw + 1:H + 2 * sw + 2 * nc:2] = p
    return recs


def _sample_dicts(recs, tr, sm):
    from galvanise_zero_b200.engine import PuctEngine
    ns = types.SimpleNamespace(n_bases=tr.num_bases, state_words=sm.state_words, role_sizes=list(tr.policy_dist_count))
    return PuctEngine.records_to_samples(ns, recs)


@pytest.mark.parametrize('game', ['reversi', 'breakthrough', 'connect6'])
def test_oracle_rows_equal_the_python_transformer_on_decoded_samples(game):
    gi, tr, ct, sm = _setup(game)
    recs = synthetic_records(tr, sm, 48, seed=5)
    enc = RefEncoder.from_c_transformer(ct)
    planes, pols, vals, bad = tref.records_to_training(recs, sm.state_words, tr.num_bases, tr.policy_dist_count,
                                                       tr.num_rewards, enc)
    assert bad == 0
    for i, d in enumerate(_sample_dicts(recs, tr, sm)):
        fs = [max(0.0, min(1.0, v)) for v in d['final_score']]                  # datacache.py:716-718
        assert np.array_equal(planes[i], tr.state_to_channels(d['state'], d['prev_states']).reshape(-1))
        for ri in range(tr.role_count):
            assert np.array_equal(pols[ri][i], tr.policy_to_array(d['policies'][ri], ri))
        assert np.array_equal(vals[i], tr.value_to_array(fs))
        tr.check_sample(datadesc.Sample(state=d['state'], prev_states=d['prev_states'], policies=d['policies'],
                                        final_score=fs))


@pytest.mark.parametrize('game', ['reversi', 'connect6'])
def test_samples_to_records_is_the_inverse_of_records_to_samples(game, tmp_path):
    """Sample objects — also after a trip through a generation file in the reference's format, states as base64
    strings — pack back into the records they came from (the fields the array builder reads), so a file written by
    the reference can feed the device path."""
    from galvanise_zero_b200 import gendata
    from galvanise_zero_b200.datacache import samples_to_records
    gi, tr, ct, sm = _setup(game)
    recs = synthetic_records(tr, sm, 40, seed=77)
    f = recs.view(np.float32)
    for i in range(len(recs)):                       # (file floats are rounded to 5 decimals: use values that survive)
        for r in range(2):
            f[i, tref.SR_FINAL_SCORE + r] = np.float32(round(float(f[i, tref.SR_FINAL_SCORE + r]), 5))
    dicts = _sample_dicts(recs, tr, sm)
    back = samples_to_records(dicts, tr, sm.state_words, recs.shape[1])
    used = [tref.SR_LEAD, tref.SR_N_PREV, tref.SR_N_CHILDREN, tref.SR_DEPTH] + list(range(tref.SR_FINAL_SCORE, tref.SR_FINAL_SCORE + 2))
    assert np.array_equal(back[:, used], recs[:, used]) and np.array_equal(back[:, H:], recs[:, H:])
    for i in range(len(recs)):
        o = 1 - recs[i, tref.SR_LEAD]
        assert back[i, tref.SR_OTHER_MOVES + o] == recs[i, tref.SR_OTHER_MOVES + o]
    samples = [gendata.encode_sample_states(datadesc.Sample(**d)) for d in dicts]
    path = str(tmp_path / gendata.gendata_filename(game, 7))
    gendata.write_gendata(path, game, 'gen_x', samples)
    gen = gendata.read_gendata(path)
    again = samples_to_records(gen.samples, tr, sm.state_words, recs.shape[1])
    st = slice(H, H + 2 * sm.state_words)
    assert np.array_equal(again[:, st], recs[:, st]) and np.array_equal(again[:, used[:3]], recs[:, used[:3]])
    enc = RefEncoder.from_c_transformer(ct)
    a = tref.records_to_training(again, sm.state_words, tr.num_bases, tr.policy_dist_count, tr.num_rewards, enc)
    b = tref.records_to_training(recs, sm.state_words, tr.num_bases, tr.policy_dist_count, tr.num_rewards, enc)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    assert all(np.allclose(x, y, atol=1e-5) for x, y in zip(a[1], b[1]))        # (probabilities rounded by the file format)


def test_host_value_policy_and_check_functions():
    gi, tr, ct, sm = _setup('reversi')                       # draw head (data/reversi_8x8/generations/*.json)
    assert tr.num_rewards == 3 and bases.GdlBasesTransformer_Draws is bases.GdlBasesTransformer
    assert tr.value_to_array([0.5, 0.5]).tolist() == [0.0, 0.0, 1.0]
    assert tr.value_to_array([0.495, 0.505]).tolist() == [0.0, 0.0, 1.0]
    assert tr.value_to_array([1.0, 0.0]).tolist() == [1.0, 0.0, 0.0]
    with pytest.raises(AssertionError):
        tr.value_to_array([0.8, 0.8])
    gi2, tr2, _, _ = _setup('breakthrough')
    assert tr2.num_rewards == 2 and tr2.value_to_array([0.25, 0.75]).dtype == np.float32
    arr = tr.policy_to_array([(3, 0.25), (64, 0.75)], 1)
    assert arr.shape == (65,) and arr[3] == np.float32(0.25) and arr[64] == np.float32(0.75) and arr.sum() == 1.0
    good = datadesc.Sample(state=[0] * tr.num_bases, policies=[[(0, 1.0)], [(1, 0.5), (2, 0.5)]], final_score=[1.0, 0.0])
    assert tr.check_sample(good) is good
    for policies in ([[(0, 1.0)], [(1, 0.5), (2, 0.3)]], [[(0, 1.2)], [(1, 1.0)]]):
        with pytest.raises(AssertionError):
            tr.check_sample(datadesc.Sample(state=[0] * tr.num_bases, policies=policies, final_score=[1.0, 0.0]))
    assert tr.get_symmetries_desc().do_reflection and tr.get_symmetries_desc().do_rotations_90
    assert _setup('chess_15d')[1].get_symmetries_desc() is None


def test_oracle_augmentation_follows_the_reference_control_flow():
    """`seen` starts empty: the identity entry takes no draw and is always yielded; duplicates (a symmetric
    position) are dropped and do not enter `seen`; an entry skipped by its draw does not enter `seen` either."""
    ident = np.arange(4)
    swap = np.array([1, 0, 3, 2])
    sym_state = [1, 1, 0, 0]                 # invariant under `swap`
    asym_state = [1, 0, 0, 0]
    rows = tref.augmentation_rows([sym_state, asym_state, asym_state], [ident, swap, swap],
                                  draws=[[0.1, 0.1], [0.1, 0.1], [0.9, 0.1]], pct=0.5)
    assert rows == [(0, 0), (1, 0), (1, 1), (2, 0), (2, 2)]


# ------------------------------------------------------------------------------------------------ warp emulation
def _warpsim(recs, sw, n_bases, widths, n_values, index=None, sym=None, base_src=None, action_dest=None, grid=2):
    """rec_unpack_kernel compiled for the CPU (tests/hostsim/warpsim_training.cpp), one std::thread per lane."""
    import ctypes
    import os
    import subprocess
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'hostsim')
    subprocess.check_call(['make', '-s', '-C', here, 'libwarpsim.so'])
    L = ctypes.CDLL(os.path.join(here, 'libwarpsim.so'))
    vp, ci, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    L.ws_records_to_training.argtypes = [i64, i64, ci, ci, ci, ci, vp, ci, vp, vp, vp, ci, vp, vp, vp, vp, vp, vp, vp,
                                         vp, ci]
    recs = np.ascontiguousarray(recs, dtype=np.int32)
    n = len(recs) if index is None else len(index)
    R = len(widths)
    pols = [np.full((n, w), 7.0, dtype=np.float32) for w in widths]
    vals = np.full((n, n_values), 7.0, dtype=np.float32)
    states = np.full((n, sw), 0xdeadbeef, dtype=np.uint32)
    prev = np.full((n, sw), 0xdeadbeef, dtype=np.uint32)
    cnt = np.full(n, 9, dtype=np.uint8)
    bad = np.zeros(1, dtype=np.int32)
    w_arr = (ci * R)(*widths)
    p_arr = (vp * R)(*[p.ctypes.data for p in pols])
    keep = []
    a_arr = None
    if sym is not None:
        keep = [np.ascontiguousarray(a, dtype=np.int32) for a in action_dest]
        a_arr = (vp * R)(*[a.ctypes.data for a in keep])
        base_src = np.ascontiguousarray(base_src, dtype=np.int32)
        sym = np.ascontiguousarray(sym, dtype=np.int32)
    if index is not None:
        index = np.ascontiguousarray(index, dtype=np.int32)
    L.ws_records_to_training(n, len(recs), recs.shape[1], sw, n_bases, R, w_arr, n_values, recs.ctypes.data,
                             index.ctypes.data if index is not None else None,
                             sym.ctypes.data if sym is not None else None,
                             len(base_src) if sym is not None else 0,
                             base_src.ctypes.data if sym is not None else None, a_arr, p_arr, vals.ctypes.data,
                             states.ctypes.data, prev.ctypes.data, cnt.ctypes.data, bad.ctypes.data, grid)
    return states, prev, cnt, pols, vals, int(bad[0])


class _RowsOnly(object):
    """Stands where the oracle encoder goes: keeps the state rows the oracle would encode."""

    def encode(self, states, prevs, cnt):
        self.rows = (states.copy(), prevs[:, 0].copy(), cnt.copy())
        return np.zeros((len(states), 1), dtype=np.float32)


@pytest.mark.timeout(300)
@pytest.mark.parametrize('game,n', [('reversi', 70), ('breakthrough', 40), ('hexLG13', 9)])
def test_kernel_source_under_warp_emulation_matches_oracle(game, n):
    """The CUDA kernel's own source, compiled for the CPU with 32 threads per warp: state rows (translated),
    dense policy rows, value rows and the check_sample count equal the oracle's, with and without symmetries and
    with an explicit row index (repeats, an index outside the batch)."""
    gi, tr, ct, sm = _setup(game)
    sw, nb, widths, nv = sm.state_words, tr.num_bases, list(tr.policy_dist_count), tr.num_rewards
    corrupt = {1, n - 2}
    recs = synthetic_records(tr, sm, n, seed=31, corrupt=corrupt)
    recs[3, tref.SR_OTHER_MOVES + (1 - recs[3, tref.SR_LEAD])] = -4
    recs[4, tref.SR_N_CHILDREN] = 10 ** 6                                  # more pairs than a record holds
    keepr = _RowsOnly()
    ref4 = recs.copy()
    ref4[4, tref.SR_N_CHILDREN] = 0                                        # (the kernel treats it as an empty policy)
    _, pols, vals, bad = tref.records_to_training(ref4, sw, nb, widths, nv, keepr)
    st, pv, cnt, gp, gv, gbad = _warpsim(recs, sw, nb, widths, nv)
    assert gbad == bad == len(corrupt) + 2
    assert np.array_equal(st, keepr.rows[0]) and np.array_equal(cnt, keepr.rows[2])
    assert np.array_equal(pv[cnt == 1], keepr.rows[1][cnt == 1]) and not pv[cnt == 0].any()
    assert all(np.array_equal(a, b) for a, b in zip(gp, pols)) and np.array_equal(gv, vals)
    # every record under every symmetry of the game, rows out of order, one index outside the batch
    desc = gamedesc.get_symmetries_desc(game)
    t = symmetry.create_translator(gi, getattr(gamedesc.Games(), game)(), desc)
    pres = list(symmetry.Prescription(desc))
    base_dest = [t.base_permutation(r, k) for r, k in pres]
    action_dest = [np.stack([t.action_permutation(ri, r, k) for r, k in pres]) for ri in range(len(widths))]
    base_src = np.empty((len(pres), nb), dtype=np.int32)
    for k, d in enumerate(base_dest):
        base_src[k, d] = np.arange(nb)
    good = [i for i in range(n) if i not in corrupt and i not in (3, 4)]
    rows = [(i, k) for k in range(len(pres)) for i in good[::-1][:12]]
    _, pols, vals, bad = tref.records_to_training(recs, sw, nb, widths, nv, keepr, rows=rows, base_dest=base_dest,
                                                  action_dest=[[a[k] for k in range(len(pres))] for a in action_dest])
    index = [i for i, k in rows] + [n + 5]
    syms = [k for i, k in rows] + [0]
    st, pv, cnt, gp, gv, gbad = _warpsim(recs, sw, nb, widths, nv, index=index, sym=syms, base_src=base_src,
                                         action_dest=action_dest, grid=3)
    m = len(rows)
    assert bad == 0 and gbad == 1
    assert np.array_equal(st[:m], keepr.rows[0]) and np.array_equal(cnt[:m], keepr.rows[2])
    assert np.array_equal(pv[:m][cnt[:m] == 1], keepr.rows[1][cnt[:m] == 1])
    assert all(np.array_equal(a[:m], b) for a, b in zip(gp, pols)) and np.array_equal(gv[:m], vals)
    assert not st[m].any() and not gv[m].any() and all(not a[m].any() for a in gp)


# ------------------------------------------------------------------------------------------------ GPU
def _batcher(tr, ct, sm, **kw):
    from galvanise_zero_b200.datacache import TrainingBatcher
    return TrainingBatcher(tr, ct, sm.state_words, device=0, **kw)


def _compare(out, planes, pols, vals):
    import torch
    got = out['planes'].reshape(planes.shape[0], -1)
    if got.dtype == torch.float32:
        assert np.array_equal(got.cpu().numpy(), planes)
    else:
        assert torch.equal(got.cpu(), torch.from_numpy(planes).to(torch.bfloat16))
    for ri in range(len(pols)):
        assert np.array_equal(out['policies'][ri].cpu().numpy(), pols[ri])
    assert np.array_equal(out['value'].cpu().numpy(), vals)


@pytest.mark.gpu
@pytest.mark.parametrize('game', ['reversi', 'breakthrough', 'connect6', 'amazons_10x10', 'chess_15d', 'hexLG13'])
def test_device_training_arrays_bit_exact(game):
    import torch
    gi, tr, ct, sm = _setup(game)
    n = 61 if game in ('hexLG13', 'chess_15d') else 333
    corrupt = {2, 17, n - 1}
    recs = synthetic_records(tr, sm, n, seed=9, corrupt=corrupt)
    recs[5, tref.SR_OTHER_MOVES + (1 - recs[5, tref.SR_LEAD])] = 10 ** 6     # an action index outside the policy
    enc = RefEncoder.from_c_transformer(ct)
    planes, pols, vals, bad = tref.records_to_training(recs, sm.state_words, tr.num_bases, tr.policy_dist_count,
                                                       tr.num_rewards, enc)
    assert bad == len(corrupt) + 1                       # ((1.02, -0.02) clamps to (1, 0): fine for the draw head)
    b = _batcher(tr, ct, sm)
    d = torch.from_numpy(recs).cuda()
    out = b.arrays(d, check=True)
    assert out['bad'] == bad and out['planes'].shape == (n, tr.num_channels, tr.num_cols, tr.num_rows)
    _compare(out, planes, pols, vals)
    _compare(b.arrays(d, plane_dtype=torch.bfloat16), planes, pols, vals)
    # an explicit row index: rows in any order, records repeated, an index outside the batch is a zero row
    idx = np.array([n - 1, 0, 0, 7, n + 3], dtype=np.int32)
    o2 = b.arrays(d, index=torch.from_numpy(idx).cuda(), check=True)
    keep = idx[:4]
    _compare(dict(planes=o2['planes'][:4], policies=[p[:4] for p in o2['policies']], value=o2['value'][:4]),
             planes[keep], [p[keep] for p in pols], vals[keep])
    assert o2['bad'] == 1 + sum(1 for i in keep if i in corrupt) and float(o2['value'][4].abs().sum()) == 0.0
    assert all(float(p[4].abs().sum()) == 0.0 for p in o2['policies'])


@pytest.mark.gpu
@pytest.mark.parametrize('game,pct', [('reversi', 0.6), ('breakthrough', 1.0), ('hexLG13', 0.5), ('connect6', 0.3)])
def test_device_augmentation_is_the_references(game, pct):
    """Rows chosen like augment_data (same draws in the same order, duplicates of symmetric positions dropped) and
    every row translated like Translator.translate_basestate / translate_action."""
    import torch
    gi, tr, ct, sm = _setup(game)
    n = 24 if game == 'hexLG13' else 120
    recs = synthetic_records(tr, sm, n, seed=21)
    # symmetric positions: the initial state (its mirror images coincide for these games)
    init, _ = _states(sm, 1, 1, 3)
    for i in (0, 3):
        recs[i, H:H + sm.state_words] = init[0].view(np.int32)
        recs[i, tref.SR_N_PREV] = 0
    desc = gamedesc.get_symmetries_desc(game)
    t = symmetry.create_translator(gi, getattr(gamedesc.Games(), game)(), desc)
    pres = list(symmetry.Prescription(desc))
    base_dest = [t.base_permutation(r, k) for r, k in pres]
    action_dest = [[t.action_permutation(ri, r, k) for r, k in pres] for ri in range(tr.role_count)]
    b = _batcher(tr, ct, sm, do_augment_data=True, data_augment_pct=pct, rng=random.Random(1234))
    rng = random.Random(1234)
    draws = [[rng.random() for _ in range(len(pres) - 1)] for _ in range(n)]
    states = [tref.decode_record(r, sm.state_words, tr.num_bases, tr.role_count)[0] for r in recs]
    rows = tref.augmentation_rows(states, base_dest, draws, pct)
    assert n <= len(rows) < n * len(pres)                # (the symmetric positions lose their duplicates)
    out = b.arrays(torch.from_numpy(recs).cuda(), check=True)
    assert list(zip(out['index'].cpu().tolist(), out['sym'].cpu().tolist())) == rows
    enc = RefEncoder.from_c_transformer(ct)
    planes, pols, vals, bad = tref.records_to_training(recs, sm.state_words, tr.num_bases, tr.policy_dist_count,
                                                       tr.num_rewards, enc, rows=rows, base_dest=base_dest,
                                                       action_dest=action_dest)
    assert out['bad'] == bad
    _compare(out, planes, pols, vals)


@pytest.mark.gpu
def test_device_training_arrays_from_self_play_records():
    """Records the engine itself produced (breakthroughSmall self-play with the synthetic network), gathered the
    way a generation is, through the batcher: equal to the oracle, every record passes check_sample."""
    import torch
    from galvanise_zero_b200.defs import confs
    from galvanise_zero_b200.engine import PuctEngine
    game = 'breakthroughSmall'
    sp = confs.selfplay_config_template(evals_per_move=40)
    eng = PuctEngine(game, sp, n_games=256, mode=0, node_cap=2048, child_cap=1 << 15, plane_dtype=torch.float32)
    got = []
    for _ in range(200):
        eng.run_synthetic(64)
        r = eng.fetch_sample_records()
        if len(r):
            got.append(r.copy())
        if sum(len(x) for x in got) >= 300:
            break
    recs = np.concatenate(got)
    assert len(recs) >= 300 and eng.totals()['errors'] == 0
    tr, ct = eng.transformer, eng.c_transformer
    sw = eng.state_words
    enc = RefEncoder.from_c_transformer(ct)
    planes, pols, vals, bad = tref.records_to_training(recs, sw, tr.num_bases, tr.policy_dist_count, tr.num_rewards, enc)
    assert bad == 0
    from galvanise_zero_b200.datacache import TrainingBatcher
    b = TrainingBatcher(tr, ct, sw, device=0)
    out = b.arrays(torch.from_numpy(np.ascontiguousarray(recs)).cuda(), check=True)
    assert out['bad'] == 0
    _compare(out, planes, pols, vals)
    assert np.allclose(out['policies'][0].sum(1).cpu().numpy(), 1.0, atol=1e-3)
