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
