"""GPU: the reference-shaped Python surface (ggplib-like state machine, Supervisor.poll with numpy
buffers, fetch_samples dicts) behaves like the reference's and agrees with the oracles."""
import attr
import numpy as np
import pytest
import torch

from galvanise_zero_b200 import bases, cppinterface, interface, lookup, tables
from galvanise_zero_b200.defs import confs, datadesc
from golden_util import FixtureMap, load_fixture

pytestmark = pytest.mark.gpu


def test_ggplib_shaped_state_machine_follows_golden_trajectory():
    game = 'breakthroughSmall'
    gi = lookup.by_name(game)
    sm = gi.get_sm()
    fx = load_fixture(game)
    fm = FixtureMap(fx, gi.propnet)
    plies = fx['games'][0]
    bs = sm.get_initial_state()
    assert gi.model.bases[0].startswith('(true ') and gi.model.actions[0][0].startswith('(does ')
    for k, rec in enumerate(plies[:25]):
        assert bs.to_list() == [int(x) for x in fm.state_bits(rec['state'])]
        sm.update_bases(bs)
        assert sm.is_terminal() == bool(rec['terminal'])
        for r in range(2):
            assert sm.get_legal_state(r).to_list() == fm.legal_set(r, rec['legal'][r])
        if rec['move'] is None:
            break
        jm = sm.get_joint_move()
        for r in range(2):
            jm.set(r, fm.move(r, rec['move'][r]))
        nxt = sm.new_base_state()
        sm.next_state(jm, nxt)
        bs = nxt
    other = sm.new_base_state()
    other.from_list(bs.to_list())
    assert other == bs and other.hash_code() == bs.hash_code()


def test_supervisor_numpy_poll_contract_and_samples():
    game = 'breakthroughSmall'
    gi = lookup.by_name(game)
    tr = bases.GdlBasesTransformer(gi, datadesc.generation_for(game))
    ct = cppinterface.create_c_transformer(tr)
    G = 64
    sup = interface.Supervisor(game, ct, G, 'test')
    conf = confs.selfplay_config_template(evals_per_move=30)
    sup.start_self_play(0, attr.asdict(conf))
    with pytest.raises(interface.GGPZeroInterfaceError):
        sup.poll(3, [])                                   # first poll must prime with n == 0
    planes = sup.poll(0, [])
    total = ct.total_size()
    rng = np.random.RandomState(0)
    samples = []
    for it in range(1500):
        assert planes is not None and planes.dtype == np.float32 and planes.ndim == 1
        n = planes.size // total
        assert 0 < n <= G and planes.size == n * total
        x = planes.reshape(n, tr.num_channels, tr.num_cols, tr.num_rows)
        assert set(np.unique(x)) <= {0.0, 1.0}
        pols = [np.ascontiguousarray(rng.rand(n, c).astype(np.float32)) for c in tr.policy_dist_count]
        v = rng.rand(n, 1).astype(np.float32)
        val = np.ascontiguousarray(np.concatenate([v, 1 - v], axis=1))
        if it == 5:
            with pytest.raises(interface.GGPZeroInterfaceError):
                sup.poll(n + 1, pols + [val])             # wrong prediction count is rejected
        planes = sup.poll(n, pols + [val])
        got = sup.fetch_samples()
        if got:
            samples.extend(got)
        if len(samples) > 40:
            break
    assert len(samples) > 40
    keys = {'state', 'prev_states', 'policies', 'final_score', 'depth', 'game_length', 'match_identifier',
            'has_resigned', 'resign_false_positive', 'starting_sample_depth', 'resultant_puct_score',
            'resultant_puct_visits'}
    for s in samples:
        assert set(s) == keys                             # datadesc.Sample fields
        datadesc.Sample(**s)
        assert len(s['state']) == tr.num_bases and len(s['policies']) == 2
        lead = max(s['policies'], key=len)
        assert abs(sum(p for _, p in lead) - 1.0) < 1e-3
        assert len(s['final_score']) == 2 and s['depth'] <= s['game_length']


class _SynthNN(object):
    """nn object for the pollers: ``gdl_bases_transformer`` + ``get_model().predict_on_batch`` that
    evaluates the shared synthetic network (include/gz_synth.h) on host planes."""

    def __init__(self, transformer, seed=1):
        import ctypes
        from galvanise_zero_b200 import lib as gzlib
        self.gdl_bases_transformer = transformer
        self.seed = seed
        self.L = gzlib.lib()
        self.ctypes = ctypes
        self.calls = 0

    def get_model(self):
        return self

    def predict_on_batch(self, planes):
        ct = self.ctypes
        tr = self.gdl_bases_transformer
        n = planes.shape[0]
        total = tr.num_channels * tr.channel_size
        x = torch.from_numpy(np.ascontiguousarray(planes, dtype=np.float32).reshape(n, total)).cuda()
        sizes = list(tr.policy_dist_count)
        pol = torch.zeros((n, sum(sizes)), dtype=torch.float32, device='cuda')
        val = torch.zeros((n, tr.num_rewards), dtype=torch.float32, device='cuda')
        vp = ct.c_void_p
        self.L.gz_synth_nn.argtypes = [vp, ct.c_int, ct.c_int64, ct.c_int, ct.c_uint64, ct.c_int,
                                       ct.POINTER(ct.c_int), ct.c_int, vp, vp, vp]
        rc = self.L.gz_synth_nn(x.data_ptr(), 0, n, total, self.seed, len(sizes), (ct.c_int * len(sizes))(*sizes),
                                tr.num_rewards, pol.data_ptr(), val.data_ptr(), None)
        assert rc == 0
        torch.cuda.synchronize()
        self.calls += 1
        pol, val = pol.cpu().numpy(), val.cpu().numpy()
        out, off = [], 0
        for c in sizes:
            out.append(np.ascontiguousarray(pol[:, off:off + c]))
            off += c
        return out + [np.ascontiguousarray(val)]


@pytest.mark.parametrize('game,evals,moves,batch', [('reversi', 150, 12, 1), ('breakthroughSmall', 200, 10, 1),
                                                    ('breakthroughSmall', 2500, 4, 1),     # variation lines several plies deep
                                                    ('reversi', 300, 10, 32), ('breakthroughSmall', 400, 12, 8)])
def test_play_poller_matches_reference_search(game, evals, moves, batch):
    """PUCTPlayer's call sequence (src/ggpzero/player/puctplayer.py:60-95) through PlayPoller /
    ggpzero_interface.Player against the reference's PuctEvaluator (oracle/_ref), same synthetic
    network: identical root visit counts, choices and reported scores, move after move."""
    from oracle import gzref
    if not gzref.available():
        pytest.skip('oracle/_ref missing')
    gi = lookup.by_name(game)
    tr = bases.GdlBasesTransformer(gi, datadesc.generation_for(game))
    # batch > 1: the reference's intra-tree workers (PUCTEvaluatorConfig.batch_size, default 32) — every poll hands
    # out up to `batch` leaves of the one tree (evaluator.cpp:722-742,1050-1079)
    conf = confs.base_puct_config(dirichlet_noise_pct=-1, choose='choose_top_visits', batch_size=batch)
    nn = _SynthNN(tr)
    sm = gi.get_sm()
    # (the in-flight discount of select draws random numbers: same generator seed as the reference harness)
    poller = cppinterface.PlayPoller(sm, nn, conf, seed=0x5EED0000)
    ref = gzref.RefEngine(tables.blob_by_name(game), poller.c_transformer).search(conf, evals, n_moves=moves)
    assert len(ref) > 3

    poller.player_reset(0)
    assert poller.player_get_move(0) == (-1, -1.0, -1)
    state = sm.get_initial_state()
    for k, r in enumerate(ref):
        poller.player_move(state, evals, -1.0)
        poller.poll_loop()
        lead = r['lead_role']
        mv, score, nodes = poller.player_get_move(lead)
        snap = poller.c_player.last_search()
        assert np.array_equal(snap['moves'], r['moves'])
        assert np.array_equal(snap['traversals'], r['traversals']), 'visit counts differ at move %d' % k
        assert snap['choice'] == r['choice']
        assert mv == int(r['moves'][r['choice']])
        assert abs(score - float(r['scores'][r['choice']])) < 1e-6
        assert 0.9 * r['nodes'] <= nodes <= r['nodes']   # the reference also counts nodes awaiting release
        other, other_score, _ = poller.player_get_move(1 - lead)
        # the probability reported to the role that is NOT on move is that role's own score of the chosen child
        # (player.cpp:105-119; PUCTPlayer passes our_role_index whoever is to move)
        assert abs(other_score - float(r['choice_scores'][1 - lead])) < 1e-6
        assert abs(score - float(r['choice_scores'][lead])) < 1e-6
        sm.update_bases(state)
        ls = sm.get_legal_state(1 - lead)
        assert ls.get_count() == 1 and ls.get_legal(0) == other       # the non-lead role's only move
        dbg = poller.player_tree_debug(3)
        assert len(dbg) >= 1 and dbg[0][1] == lead
        # Player::treeDebugInfo (player.cpp:137-153): scores, moves and the variation lines are the reference's
        assert len(dbg) == len(r['tree_debug'])
        for mine_ln, ref_ln in zip(dbg, r['tree_debug']):
            assert abs(mine_ln[0] - ref_ln[0]) < 1e-6 and mine_ln[1:] == ref_ln[1:]
        jm = sm.get_joint_move()
        jm.set(lead, mv)
        jm.set(1 - lead, other)
        poller.player_apply_move(jm)
        poller.poll_loop()
        nxt = sm.new_base_state()
        sm.next_state(jm, nxt)
        state = nxt
    assert nn.calls > (evals if batch == 1 else evals // batch)


def test_player_balance_moves_matches_reference():
    """PUCTPlayer.balance_moves (src/ggpzero/player/puctplayer.py:97-98) -> Player::balanceNode ->
    PuctEvaluator::balanceFirstMoves (evaluator.cpp:282-339) through the poll loop: the root's traversal counts after
    balancing the first three moves are the reference's, and so is every later search on the balanced tree."""
    from oracle import gzref
    if not gzref.available():
        pytest.skip('oracle/_ref missing')
    game, evals, moves, batch, k = 'reversi', 300, 5, 8, 3
    gi = lookup.by_name(game)
    tr = bases.GdlBasesTransformer(gi, datadesc.generation_for(game))
    conf = confs.base_puct_config(dirichlet_noise_pct=-1, choose='choose_top_visits', batch_size=batch)
    nn = _SynthNN(tr)
    sm = gi.get_sm()
    poller = cppinterface.PlayPoller(sm, nn, conf, seed=0x5EED0000)
    eng = gzref.RefEngine(tables.blob_by_name(game), poller.c_transformer)
    eng.set_balance(k)
    try:
        ref = eng.search(conf, evals, n_moves=moves)
    finally:
        eng.set_balance(0)
    poller.player_reset(0)
    state = sm.get_initial_state()
    grew = 0
    for r in ref:
        poller.player_move(state, evals, -1.0)
        poller.poll_loop()
        lead = r['lead_role']
        mv, _, _ = poller.player_get_move(lead)
        assert np.array_equal(poller.c_player.last_search()['traversals'], r['traversals'])
        poller.player_balance_moves(k)
        poller.poll_loop()
        snap = poller.c_player.engine.snapshot(0)
        assert np.array_equal(snap['traversals'], r['balanced_traversals'])
        grew += int(snap['traversals'].sum() > r['traversals'].sum())
        other, _, _ = poller.player_get_move(1 - lead)
        jm = sm.get_joint_move()
        jm.set(lead, mv)
        jm.set(1 - lead, other)
        poller.player_apply_move(jm)
        poller.poll_loop()
        nxt = sm.new_base_state()
        sm.next_state(jm, nxt)
        state = nxt
    assert grew >= 2


def test_supervisor_poller_wrapper_collects_samples():
    """cppinterface.Supervisor (src/ggpzero/util/cppinterface.py:185-226) driven with poll_loop and
    a callback, as Worker.on_request_samples does (src/ggpzero/distributed/worker.py:176-191)."""
    game = 'breakthroughSmall'
    gi = lookup.by_name(game)
    tr = bases.GdlBasesTransformer(gi, datadesc.generation_for(game))
    nn = _SynthNN(tr)
    sup = cppinterface.Supervisor(gi.get_sm(), nn, batch_size=64, identifier='w0')
    sup.start_self_play(confs.selfplay_config_template(evals_per_move=20), 0)
    got = []

    def cb():
        got.extend(sup.fetch_samples())
        return len(got) >= 30
    sup.poll_loop(cb=cb, do_stats=True)
    assert len(got) >= 30 and all(isinstance(s, datadesc.Sample) for s in got)
    assert sup.total_predictions > 0 and sup.acc_time_polling > 0
    sup.clear_unique_states()
    # resumable: a second loop keeps producing (src/test/cpp/test_interface.py:166-170)
    n0 = len(got)
    sup.poll_loop(cb=lambda: bool(got.extend(sup.fetch_samples())) or len(got) >= n0 + 10)
    assert len(got) >= n0 + 10


def test_worker_rehost_request_samples_and_generation_file(tmp_path):
    """SelfPlayWorker (src/ggpzero/distributed/worker.py:138-206 re-hosted): worker-mode supervisor (two
    managers, double-buffered), samples with base64 states, written and re-read as a generation file."""
    from galvanise_zero_b200 import gendata, worker
    game = 'breakthroughSmall'
    gi = lookup.by_name(game)
    tr = bases.GdlBasesTransformer(gi, datadesc.generation_for(game))
    w = worker.SelfPlayWorker(gi.get_sm(), _SynthNN(tr), self_play_batch_size=32, min_num_samples=25,
                              identifier='wk')
    w.configure_self_play(confs.selfplay_config_template(evals_per_move=20))
    assert len(w.supervisor.c_supervisor.managers) == 2
    samples = w.on_request_samples()
    assert len(samples) > 25 and w.last_stats['predictions'] > 0
    assert all(isinstance(s.state, str) for s in samples)
    idents = {s.match_identifier.split('_')[1] for s in samples}
    assert idents <= {'m0', 'm1'}                      # samples come from both managers
    path = str(tmp_path / gendata.gendata_filename(game, 1))
    gendata.write_gendata(path, game, 'b200_1', samples)
    back = gendata.read_gendata(path)
    assert back.num_samples == len(samples)
    st = gendata.decode_state(back.samples[0].state)
    assert len(st) >= tr.num_bases and set(st) <= {0, 1}
    more = w.on_request_samples()                      # resumable
    assert len(more) > 25


def test_fetch_with_small_buffer_keeps_the_rest_and_counts_nothing_as_dropped():
    """gz_puct_fetch_samples hands out at most `cap` records and keeps the others for the next call (the
    reference's sample vector is unbounded, supervisor.cpp:170-173)."""
    from galvanise_zero_b200.engine import PuctEngine
    sp = confs.selfplay_config_template(evals_per_move=20)
    eng = PuctEngine('breakthroughSmall', sp, n_games=64, mode=0, node_cap=1024, child_cap=1024 * 16)
    for _ in range(400):
        eng.run_synthetic(64)
        if eng.totals()['ready'] >= 40:
            break
    t = eng.totals()
    ready = int(t['ready'])
    assert ready >= 40 and t['dropped_samples'] == 0
    first = eng.fetch_sample_records(cap=7)
    assert len(first) == 7
    assert int(eng.totals()['ready']) == ready - 7
    rest = eng.fetch_sample_records()
    assert len(rest) >= ready - 7
    keys = {(int(r[3]), int(r[4]), int(r[1])) for r in np.concatenate([first, rest])}
    assert len(keys) == len(first) + len(rest)            # nothing duplicated, nothing lost
    assert all(int(r[0]) == 1 for r in rest)


def test_full_ready_buffer_is_counted_and_reported():
    from galvanise_zero_b200.engine import PuctEngine
    sp = confs.selfplay_config_template(evals_per_move=16)
    eng = PuctEngine('breakthroughSmall', sp, n_games=64, mode=0, node_cap=1024, child_cap=1024 * 16, ready_cap=8)
    for _ in range(40):
        eng.run_synthetic(64)
    t = eng.totals()
    assert t['ready'] == 8 and t['dropped_samples'] > 0 and t['samples'] == t['dropped_samples'] + 8


def test_engine_rejects_more_than_one_previous_state():
    from galvanise_zero_b200.engine import PuctEngine
    from galvanise_zero_b200 import lib as gzlib
    gen = datadesc.generation_for('reversi', num_previous_states=2)
    with pytest.raises(gzlib.GzError):
        PuctEngine('reversi', confs.selfplay_config_template(), n_games=4, generation=gen)


def test_game_over_capacity_restarts_instead_of_halting():
    """A game that exceeds a per-game capacity (here: 3 pending samples) is abandoned and counted; its slot
    starts a new game instead of halting for good."""
    from galvanise_zero_b200.engine import PuctEngine
    sp = confs.selfplay_config_template(evals_per_move=16)
    sp.oscillate_sampling_pct = 1.0
    eng = PuctEngine('breakthroughSmall', sp, n_games=32, mode=0, node_cap=1024, child_cap=1024 * 16,
                     max_pending_samples=3)
    for _ in range(30):
        eng.run_synthetic(64)
    t = eng.totals()
    assert t['halted'] == 0 and t['errors'] == 0 and t['error_restarts'] > 0 and t['games'] >= 0
