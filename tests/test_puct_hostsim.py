"""CPU: the PUCT / self-play core (csrc/gz_puct_core.h, compiled for the host by tests/hostsim) is
compared with the reference's own C++ (puct/evaluator.cpp, puct/node.cpp, selfplay.cpp, ... compiled
unmodified into oracle/_ref/libgzref.so) under an identical synthetic network and RNG streams.

Bar: EXACT equality of root visit counts per move and of every self-play sample record
(state, previous state, policy, final scores, depth, game length, resign flags) — i.e. total-variation
distance 0, well inside the stated tolerance (TV <= 0.02). Every RNG-driven decision
(oscillating sampling, temperature move choice, resign permission, root latch) is on; Dirichlet noise —
which the reference draws with libstdc++'s std::gamma_distribution<float>, restated bit for bit in
csrc/gz_libm.h — is off in the first tests and ON in the *_with_dirichlet_noise ones."""
import numpy as np
import pytest

from galvanise_zero_b200 import bases, cppinterface, lookup, tables
from galvanise_zero_b200.defs import confs, datadesc
from oracle import gzref

pytestmark = pytest.mark.skipif(not gzref.available(), reason='oracle/_ref cannot be built here')


def _setup(game):
    from hostsim_util import HostSim  # noqa: F401
    gi = lookup.by_name(game)
    tr = bases.GdlBasesTransformer(gi, datadesc.generation_for(game))
    ct = cppinterface.create_c_transformer(tr)
    eng = gzref.RefEngine(tables.blob_by_name(game), ct)
    return ct, eng


@pytest.mark.parametrize('game,evals,moves', [('reversi', 200, 12), ('breakthrough', 300, 25),
                                              ('connect6', 120, 4), ('amazons_10x10', 150, 4)])
def test_visit_counts_match_reference_search(game, evals, moves):
    from hostsim_util import HostSim
    ct, eng = _setup(game)
    conf = confs.base_puct_config(dirichlet_noise_pct=-1, choose='choose_top_visits')
    sp = confs.selfplay_config_template()
    sp.puct_config = conf
    ref = eng.search(conf, evals, n_moves=moves)
    mine = HostSim(tables.blob_by_name(game), ct, sp).search(evals, moves)
    assert len(ref) == len(mine) == moves
    for a, b in zip(ref, mine):
        assert np.array_equal(a['moves'], b['moves'])
        assert np.array_equal(a['traversals'], b['traversals'])       # TV distance == 0
        assert a['choice'] == b['choice'] and a['root_visits'] == b['root_visits']
        assert np.array_equal(a['policy_prob'], b['policy_prob'])
        assert np.array_equal(a['scores'], b['scores'])


@pytest.mark.parametrize('game,evals,moves,squash', [('reversi', 200, 14, -1), ('breakthrough', 150, 20, 0.5),
                                                     ('connect6', 100, 3, -1), ('amazons_10x10', 80, 3, -1)])
def test_visit_counts_match_with_dirichlet_noise(game, evals, moves, squash):
    """setDirichletNoise (evaluator.cpp:1227-1297) ON at every root, incl. the alpha < 1 branch (more than 10
    children: connect6 361, amazons 2176) and the policy squash: priors after noise and visit counts are the
    reference's, value for value."""
    from hostsim_util import HostSim
    ct, eng = _setup(game)
    conf = confs.base_puct_config(dirichlet_noise_pct=0.25, choose='choose_top_visits',
                                  noise_policy_squash_pct=squash, noise_policy_squash_prob=0.1)
    sp = confs.selfplay_config_template()
    sp.puct_config = conf
    ref = eng.search(conf, evals, n_moves=moves, max_children=4096)
    mine = HostSim(tables.blob_by_name(game), ct, sp).search(evals, moves)
    assert len(ref) == len(mine) == moves
    for a, b in zip(ref, mine):
        assert np.array_equal(a['policy_prob'], b['policy_prob'])     # the noised priors themselves
        assert np.array_equal(a['traversals'], b['traversals'])
        assert a['choice'] == b['choice'] and a['root_visits'] == b['root_visits']
        assert np.array_equal(a['scores'], b['scores'])


def test_selfplay_samples_match_with_dirichlet_noise():
    """The benchmarked configuration (templates.py:73-129: noise 0.25 in sampled searches, 0.15 in run-to-end)."""
    _selfplay_samples('reversi', 4, 50, 60, noise=True)


@pytest.mark.parametrize('game,evals,moves', [('breakthroughSmall', 400, 20), ('reversi', 600, 10), ('connect6', 200, 4),
                                              ('amazons_10x10', 150, 3)])
def test_transposition_lookup_matches_reference(game, evals, moves):
    """PuctConfig.lookup_transpositions (evaluator.cpp:144-173,216-239): a state reached by two move orders shares
    one node (the tree becomes a DAG), the playout that finds it ends without an evaluation, and re-rooting keeps
    every node some kept node still links to. Visit counts, node counts and the number of attached transpositions
    are the reference's."""
    from hostsim_util import HostSim
    ct, eng = _setup(game)
    conf = confs.base_puct_config(dirichlet_noise_pct=-1, choose='choose_top_visits', lookup_transpositions=True)
    sp = confs.selfplay_config_template()
    sp.puct_config = conf
    ref = eng.search(conf, evals, n_moves=moves, max_children=4096)
    hs = HostSim(tables.blob_by_name(game), ct, sp)
    mine = hs.search(evals, moves, max_children=4096)
    assert len(ref) == len(mine) == moves
    for m, (a, b) in enumerate(zip(ref, mine)):
        assert np.array_equal(a['moves'], b['moves'])
        assert np.array_equal(a['traversals'], b['traversals'])
        assert a['choice'] == b['choice'] and a['root_visits'] == b['root_visits']
        assert np.array_equal(a['scores'], b['scores'])
        assert a['nodes'] == b['nodes'] + m     # the reference keeps the chain of past roots, the arena does not
    if game in ('breakthroughSmall', 'reversi'):
        assert hs.transpositions() > 10     # the case is not vacuous (the wide games rarely transpose this shallow)


@pytest.mark.parametrize('game,evals,moves,batch,tt,noise', [
    ('breakthroughSmall', 400, 20, 8, False, -1), ('reversi', 600, 12, 32, False, -1),
    ('reversi', 300, 20, 4, False, 0.25), ('breakthrough', 500, 12, 16, True, -1),
    ('connect6', 200, 3, 32, False, -1), ('amazons_10x10', 150, 2, 8, False, -1),
    ('reversi', 60, 6, 32, False, -1)])      # <= 100 evaluations: no workers are spawned (evaluator.cpp:1057)
def test_intra_tree_workers_match_reference(game, evals, moves, batch, tt, noise):
    """PuctConfig.batch_size > 1 (a Player's default is 32, confs.py:64): playoutMain and batch_size - 1 playoutWorker
    coroutines share one tree (evaluator.cpp:722-742,1050-1079), leaves are evaluated in batches, in-flight visits
    discount select (evaluator.cpp:402-405,443-446, with their random draws) and children being expanded are
    unselectable. The reference's greenlet schedule (scheduler.cpp:131-203) decides every visit count; the restated
    run queue reproduces it: identical root statistics after every search of a game."""
    from hostsim_util import HostSim
    ct, eng = _setup(game)
    conf = confs.base_puct_config(dirichlet_noise_pct=noise, choose='choose_top_visits', batch_size=batch,
                                  lookup_transpositions=tt)
    sp = confs.selfplay_config_template()
    sp.puct_config = conf
    ref = eng.search(conf, evals, n_moves=moves, max_children=4096)
    mine = HostSim(tables.blob_by_name(game), ct, sp, ncap=1 << 14, ccap=1 << 21).search(evals, moves, max_children=4096)
    assert len(ref) == len(mine) == moves
    for a, b in zip(ref, mine):
        assert np.array_equal(a['moves'], b['moves'])
        assert np.array_equal(a['traversals'], b['traversals'])
        assert a['choice'] == b['choice'] and a['root_visits'] == b['root_visits']
        assert np.array_equal(a['policy_prob'], b['policy_prob'])
        assert np.array_equal(a['scores'], b['scores'])


@pytest.mark.parametrize('batch,tt', [(1, True), (8, False), (8, True)])
def test_backup_finalised_with_shared_nodes_matches(batch, tt):
    """MCTS-prover style finalisation (backup_finalised, evaluator.cpp:524-571) reads the children's nodes; with
    transpositions or workers a node has several parent records, so nothing may rely on a record's copy."""
    from hostsim_util import HostSim
    ct, eng = _setup('breakthroughSmall')
    conf = confs.base_puct_config(dirichlet_noise_pct=-1, backup_finalised=True, choose='choose_top_visits',
                                  batch_size=batch, lookup_transpositions=tt)
    sp = confs.selfplay_config_template()
    sp.puct_config = conf
    ref = eng.search(conf, 400, n_moves=30)
    mine = HostSim(tables.blob_by_name('breakthroughSmall'), ct, sp).search(400, 30)
    assert len(ref) == len(mine) and len(ref) > 8
    for a, b in zip(ref, mine):
        assert np.array_equal(a['traversals'], b['traversals']) and a['choice'] == b['choice']
        assert np.array_equal(a['scores'], b['scores'])


@pytest.mark.parametrize('game,evals,batch,k', [('breakthroughSmall', 300, 8, 4), ('reversi', 400, 16, 3),
                                                ('reversi', 300, 2, 5)])
def test_balance_first_moves_matches_reference(game, evals, batch, k):
    """Player::balanceNode -> PuctEvaluator::balanceFirstMoves (player.cpp:121-135, evaluator.cpp:282-339): after a
    search, the k most visited root children get batches of batch_size playouts until each has as many traversals as
    the most visited one. Same traversal counts as the reference, move after move."""
    from hostsim_util import HostSim
    ct, eng = _setup(game)
    conf = confs.base_puct_config(dirichlet_noise_pct=-1, choose='choose_top_visits', batch_size=batch)
    sp = confs.selfplay_config_template()
    sp.puct_config = conf
    eng.set_balance(k)
    try:
        ref = eng.search(conf, evals, n_moves=6, max_children=512)
    finally:
        eng.set_balance(0)
    mine = HostSim(tables.blob_by_name(game), ct, sp, ncap=1 << 14, ccap=1 << 20).search(evals, 6, balance=k)
    assert len(ref) == len(mine) == 6
    grew = 0
    for a, b in zip(ref, mine):
        assert np.array_equal(a['traversals'], b['traversals'])
        assert np.array_equal(a['balanced_traversals'], b['balanced_traversals'])
        grew += int(b['balanced_traversals'].sum() > b['traversals'].sum())
        assert b['balanced_traversals'].sum() >= b['traversals'].sum()
        assert a['choice'] == b['choice']
    assert grew >= 3                   # the case is not vacuous


def test_backup_finalised_and_temperature_choice_match():
    from hostsim_util import HostSim
    ct, eng = _setup('breakthroughSmall')
    conf = confs.base_puct_config(dirichlet_noise_pct=-1, backup_finalised=True)   # choose_temperature
    sp = confs.selfplay_config_template()
    sp.puct_config = conf
    ref = eng.search(conf, 150, n_moves=30)
    mine = HostSim(tables.blob_by_name('breakthroughSmall'), ct, sp).search(150, 30)
    assert len(ref) == len(mine) and len(ref) > 8
    for a, b in zip(ref, mine):
        assert np.array_equal(a['traversals'], b['traversals']) and a['choice'] == b['choice']


@pytest.mark.parametrize('game,G,evals,target', [('breakthrough', 6, 60, 120), ('breakthroughSmall', 8, 40, 150),
                                                 ('reversi', 4, 50, 60)])
def test_selfplay_samples_match_reference_supervisor(game, G, evals, target):
    _selfplay_samples(game, G, evals, target, noise=False)


def test_capped_batches_defer_requests_without_changing_any_search():
    """Compacted network batches (gz_puct_compact_requests / gz_puct_advance_compact): when the batch holds fewer
    rows than there are requests, the overflow keeps its request and waits one more step. A search is a deterministic
    process of its own, so four trees searched side by side end with the same visit counts whether one, two or all
    four requests are answered per step."""
    from hostsim_util import HostSim
    ct, eng = _setup('reversi')
    conf = confs.base_puct_config(dirichlet_noise_pct=0.25, choose='choose_top_visits')
    sp = confs.selfplay_config_template()
    sp.puct_config = conf
    G, evals = 4, 150
    results = []
    for cap in (0, 1, 2):
        hs = HostSim(tables.blob_by_name('reversi'), ct, sp, G=G, mode=1)
        hs.set_answer_cap(cap)
        L = hs.L
        for g in range(G):
            L.hs_cmd_reset(hs.h, g, None, 0)
        for g in range(G):
            assert L.hs_wait(hs.h, g, 100000) == 1
        for g in range(G):
            L.hs_cmd_search(hs.h, g, evals)
        for g in range(G):
            assert L.hs_wait(hs.h, g, 10000000) == 1
        per_game = []
        for g in range(G):
            moves = np.zeros(128, dtype=np.int32); trav = np.zeros(128, dtype=np.int32)
            prob = np.zeros(128, dtype=np.float32); score = np.zeros(128, dtype=np.float32)
            info = np.zeros(8, dtype=np.float32)
            k = L.hs_root_children(hs.h, g, 128, moves.ctypes.data, trav.ctypes.data, prob.ctypes.data,
                                   score.ctypes.data, info.ctypes.data)
            per_game.append((trav[:k].copy(), prob[:k].copy(), score[:k].copy(), int(info[0])))
        results.append(per_game)
    for other in results[1:]:
        for a, b in zip(results[0], other):
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3]
    # the games differ from each other (their own RNG streams draw their own root noise)
    assert not np.array_equal(results[0][0][1], results[0][1][1])


def test_selfplay_samples_match_with_transposition_lookup():
    """Whole games with shared nodes: every re-root rebuilds the table, samples stay the reference's."""
    _selfplay_samples('breakthroughSmall', 6, 60, 150, noise=False, transpositions=True)


def _selfplay_samples(game, G, evals, target, noise, transpositions=False):
    from hostsim_util import HostSim
    ct, eng = _setup(game)
    sp = confs.selfplay_config_template(evals_per_move=evals)
    sp.puct_config.lookup_transpositions = transpositions
    sp.run_to_end_puct_config.lookup_transpositions = transpositions
    if not noise:
        sp.puct_config.dirichlet_noise_pct = -1
        sp.run_to_end_puct_config.dirichlet_noise_pct = -1
    pn = tables.propnet_by_name(game)
    SW = tables.state_words(pn.n_bases)
    srec = 24 + 2 * SW + 2 * max(pn.role_sizes)
    ref = eng.selfplay(sp, G, 0, target_samples=target, max_seconds=240, sample_capacity=target * 6, srec=srec)
    recs = ref['records']
    assert len(recs) >= target
    hs = HostSim(tables.blob_by_name(game), ct, sp, G=G, mode=0)
    for _ in range(200000):
        if hs.L.hs_ready_count(hs.h) >= len(recs):
            break
        hs.step(32)
    n = hs.L.hs_ready_count(hs.h)
    mine = np.ctypeslib.as_array(hs.L.hs_ready(hs.h), shape=(n * srec,)).reshape(n, srec)

    def key(r):
        return (int(r[3]), int(r[4]), int(r[1]))
    rd = {key(r): r for r in recs}
    md = {key(r): r for r in mine}
    # the reference hands samples over in bursts; compare every sample both sides have produced
    common = sorted(set(rd) & set(md))
    assert len(common) >= 0.8 * len(rd)
    for k in common:
        a, b = rd[k], md[k]
        nc = int(a[11])
        assert np.array_equal(a[:24 + 2 * SW + 2 * nc], b[:24 + 2 * SW + 2 * nc]), k
    # and per game slot the sequence of sampled (match, depth) pairs is identical up to the shorter one
    for slot in range(G):
        ka = sorted(k for k in rd if k[0] == slot)
        kb = sorted(k for k in md if k[0] == slot)
        m = min(len(ka), len(kb))
        assert ka[:m] == kb[:m]
