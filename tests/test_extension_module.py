"""The reference's extension module, importable under its own name: ``import ggpzero_interface``
(src/cpp/ggpzero_interface.cpp:50-96) — factories, argument formats and method tables as the reference's
bindings declare them (src/cpp/pyobjects/*_impl.cpp). CPU: the surface; GPU: a self-play supervisor and a player
driven through it, checked against the ctypes host and the reference search."""
import os
import sys

import numpy as np
import pytest

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'galvanise_zero_b200', 'csrc')


def _module():
    from galvanise_zero_b200 import build
    build.build_extension()
    if CSRC not in sys.path:
        sys.path.insert(0, CSRC)
    import ggpzero_interface
    return ggpzero_interface


def test_module_surface_and_argument_formats():
    gi = _module()
    assert gi.__name__ == 'ggpzero_interface'
    for name in ('GdlBasesTransformer', 'Supervisor', 'Player', 'buf_to_tuple_reverse_bytes', 'GGPZeroInterfaceError'):
        assert hasattr(gi, name)
    assert issubclass(gi.GGPZeroInterfaceError, Exception)
    # "iiiiiO!": five ints and a LIST of policy sizes
    t = gi.GdlBasesTransformer(64, 2, 1, 1, 3, [65, 65])
    with pytest.raises(TypeError):
        gi.GdlBasesTransformer(64, 2, 1, 1, 3, (65, 65))
    with pytest.raises(TypeError):
        gi.GdlBasesTransformer(64, 2, 1, 1, 3)
    assert t.add_board_base(3, 5) is None            # "ii"; returns a properly counted None
    assert t.add_control_base(1, 0, 0.5) is None     # "iif"
    with pytest.raises(TypeError):
        t.add_control_base(1, 0)
    assert t.total_size() == 64 * (2 * 2 + 1)
    for m in ('add_board_base', 'add_control_base', 'test'):
        assert hasattr(t, m)
    # "nO!is" / "nO!O!": a pointer as an int, the transformer OBJECT, ...
    with pytest.raises(TypeError):
        gi.Supervisor(0, object(), 4, 'x')
    with pytest.raises(TypeError):
        gi.Supervisor(0, t, 4)
    with pytest.raises(TypeError):
        gi.Player(0, t, [('a', 1)])
    # bytes -> bools, most significant bit first (inverse of np.packbits)
    bits = np.array([1, 0, 0, 0, 0, 0, 0, 1, 0, 1, 1, 0, 0, 0, 0, 0], dtype=np.uint8)
    assert gi.buf_to_tuple_reverse_bytes(np.packbits(bits).tobytes()) == tuple(bool(b) for b in bits)


@pytest.mark.gpu
def test_supervisor_through_the_extension_module():
    import attr
    from galvanise_zero_b200 import bases, cppinterface, lookup
    from galvanise_zero_b200.defs import confs, datadesc
    gi = _module()
    game = 'breakthroughSmall'
    info = lookup.by_name(game)
    sm = info.get_sm()
    tr = bases.GdlBasesTransformer(info, datadesc.generation_for(game))
    ct = cppinterface.create_c_transformer(tr, module=gi)       # built through the module's own factory
    assert type(ct).__name__ == 'GdlBasesTransformerWrapper'
    sup = gi.Supervisor(sm.sm_ptr, ct, 32, 'ext')
    conf = confs.selfplay_config_template(evals_per_move=24)
    sup.start_self_play(1, attr.asdict(conf))                   # worker mode: two managers
    with pytest.raises(gi.GGPZeroInterfaceError):
        sup.poll(3, [])                                         # the first poll primes with n == 0
    planes = sup.poll(0, [])
    total = ct.total_size()
    rng = np.random.RandomState(0)
    samples = []
    for it in range(3000):
        assert isinstance(planes, np.ndarray) and planes.dtype == np.float32 and planes.ndim == 1
        n = planes.size // total
        assert 0 < n <= 32 and planes.size == n * total
        x = planes.reshape(n, tr.num_channels, tr.num_cols, tr.num_rows)
        assert set(np.unique(x)).issubset({0.0, 1.0})
        pol = [np.ascontiguousarray(rng.rand(n, w).astype(np.float32)) for w in tr.policy_dist_count]
        val = np.ascontiguousarray(rng.rand(n, 2).astype(np.float32))
        with pytest.raises(gi.GGPZeroInterfaceError):
            sup.poll(n, pol + [val.astype(np.float64)])         # float64 is refused (common.cpp:186-196)
        planes = sup.poll(n, pol + [val])
        if it % 200 == 199:
            samples.extend(sup.fetch_samples() or [])
        if len(samples) >= 40:
            break
    assert len(samples) >= 40
    s = samples[0]
    assert set(s) == {'state', 'prev_states', 'policies', 'final_score', 'depth', 'game_length', 'match_identifier',
                      'has_resigned', 'resign_false_positive', 'starting_sample_depth', 'resultant_puct_score',
                      'resultant_puct_visits'}
    assert len(s['state']) == tr.num_bases and set(s['state']) <= {0, 1}
    assert abs(sum(p for _, p in max(s['policies'], key=len)) - 1.0) < 1e-3
    assert s['match_identifier'].startswith('ext_m')
    sample = datadesc.Sample(**s)                               # the reference's attrs class takes it as is
    assert sample.depth <= sample.game_length
    t = sup.totals()
    assert t['halted'] == 0 and t['dropped_samples'] == 0
    sup.clear_unique_states()
    assert sup.add_unique_state(0) is None


@pytest.mark.gpu
def test_player_through_the_extension_module_matches_reference_search():
    """PUCTPlayer's call sequence (src/ggpzero/player/puctplayer.py:60-95) with pointers as the reference passes
    them: identical visit-count-derived choices and scores to the reference's PuctEvaluator."""
    import attr
    from galvanise_zero_b200 import bases, cppinterface, lookup, tables
    from galvanise_zero_b200.defs import confs, datadesc
    from oracle import gzref
    from test_gpu_api import _SynthNN
    if not gzref.available():
        pytest.skip('oracle/_ref missing')
    gi = _module()
    game, evals, moves = 'reversi', 120, 8
    info = lookup.by_name(game)
    sm = info.get_sm()
    tr = bases.GdlBasesTransformer(info, datadesc.generation_for(game))
    ct = cppinterface.create_c_transformer(tr, module=gi)
    conf = confs.base_puct_config(dirichlet_noise_pct=-1, choose='choose_top_visits')
    player = gi.Player(sm.sm_ptr, ct, attr.asdict(conf))
    ref = gzref.RefEngine(tables.blob_by_name(game), cppinterface.create_c_transformer(tr)).search(conf, evals, n_moves=moves)
    nn = _SynthNN(tr)

    def poll_loop():
        planes = player.poll(0, [])
        while planes is not None:
            x = planes.reshape(1, tr.num_channels, tr.num_cols, tr.num_rows)
            arrays = nn.predict_on_batch(x)
            planes = player.poll(1, arrays)

    player.player_reset(0)
    assert player.player_get_move(0) == (-1, -1.0, -1)
    state = sm.get_initial_state()
    for k, r in enumerate(ref):
        player.player_move(state.ptr, evals, -1.0)
        poll_loop()
        lead = r['lead_role']
        mv, score, nodes = player.player_get_move(lead)
        assert mv == int(r['moves'][r['choice']]), 'choice differs at move %d' % k
        assert abs(score - float(r['choice_scores'][lead])) < 1e-6
        other, other_score, _ = player.player_get_move(1 - lead)
        assert abs(other_score - float(r['choice_scores'][1 - lead])) < 1e-6
        dbg = player.player_tree_debug(3)
        assert len(dbg) >= 1 and dbg[0][1] == lead
        assert len(dbg) == len(r['tree_debug'])      # (score, lead, move, ((lead, move), ..)), player_impl.cpp:67-100
        for mine_ln, ref_ln in zip(dbg, r['tree_debug']):
            assert abs(mine_ln[0] - ref_ln[0]) < 1e-6 and tuple(mine_ln[1:]) == tuple(ref_ln[1:])
        jm = sm.get_joint_move()
        jm.set(lead, mv)
        jm.set(1 - lead, other)
        player.player_apply_move(jm.ptr)
        poll_loop()
        nxt = sm.new_base_state()
        sm.next_state(jm, nxt)
        state = nxt
    player.player_update_config(0.5, 1, False)
    player.player_balance_moves(3)


@pytest.mark.gpu
def test_player_honours_wall_clock_limits():
    """end_time (hard limit) and think_time (evaluator.cpp:794,820-826): a search with a huge evaluation budget
    returns once the clock says so."""
    import time
    import attr
    from galvanise_zero_b200 import bases, cppinterface, interface, lookup
    from galvanise_zero_b200.defs import confs, datadesc
    from test_gpu_api import _SynthNN
    game = 'reversi'
    info = lookup.by_name(game)
    sm = info.get_sm()
    tr = bases.GdlBasesTransformer(info, datadesc.generation_for(game))
    ct = cppinterface.create_c_transformer(tr)
    nn = _SynthNN(tr)

    def run(player, end_time):
        player.player_reset(0)
        player.player_move(sm.get_initial_state(), 10 ** 7, end_time)
        t0 = time.time()
        planes = player.poll(0, [])
        while planes is not None:
            planes = player.poll(1, nn.predict_on_batch(planes.reshape(1, tr.num_channels, tr.num_cols, tr.num_rows)))
        return time.time() - t0

    conf = confs.base_puct_config(dirichlet_noise_pct=-1, choose='choose_top_visits')       # think_time = -1
    p1 = interface.Player(sm, ct, attr.asdict(conf))
    dt = run(p1, time.time() + 1.0)
    assert 0.9 <= dt < 4.0
    assert p1.player_get_move(0)[0] >= 0 or p1.player_get_move(1)[0] >= 0
    conf2 = confs.base_puct_config(dirichlet_noise_pct=-1, choose='choose_top_visits', think_time=0.5,
                                   evaluation_multiplier_to_convergence=2.0, converged_visits=10 ** 9)
    p2 = interface.Player(sm, ct, attr.asdict(conf2))
    dt = run(p2, -1.0)                      # never converges: stops at think_time x multiplier = 1 s
    assert 0.9 <= dt < 4.0
