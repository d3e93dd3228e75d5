"""GPU parity: the CUDA propnet kernels, called through the C-ABI, are bit-exact with the C oracle
(oracle/propnet_ref.c) on seeded random playouts: legal masks, next states, terminal, goals."""
import numpy as np
import pytest
import torch

from galvanise_zero_b200 import lib as gzlib
from galvanise_zero_b200 import tables
from oracle.pnref import RefStateMachine

pytestmark = pytest.mark.gpu

GAMES = ['breakthrough', 'reversi', 'connect6', 'amazons_10x10', 'breakthroughSmall']


def _sm(game, **kw):
    from galvanise_zero_b200.statemachine import DeviceStateMachine
    return DeviceStateMachine(game, **kw)


def _np(t):
    a = t.cpu().numpy()
    return a.view(np.uint32) if a.dtype == np.int32 else a


@pytest.mark.parametrize('game', GAMES)
@pytest.mark.parametrize('n', [1, 33, 1000])
def test_playout_steps_bit_exact(game, n):
    dev = _sm(game)
    ref = RefStateMachine(tables.blob_by_name(game))
    assert dev.state_words == ref.state_words and dev.legal_words == ref.legal_words
    steps = {'breakthrough': 90, 'reversi': 70, 'connect6': 25, 'amazons_10x10': 30,
             'breakthroughSmall': 60}[game]
    if n == 1000:
        steps = min(steps, 40)
    seed = 0xB200 + n
    # oracle arrays
    s_r = np.zeros((n, ref.state_words), dtype=np.uint32)
    l_r = np.zeros((n, ref.legal_words), dtype=np.uint32)
    t_r = np.ones(n, dtype=np.uint8)
    g_r = np.zeros((n, ref.n_roles), dtype=np.uint8)
    m_r = np.zeros((n, ref.n_roles), dtype=np.int16)
    # device arrays
    s_d = torch.zeros((n, dev.state_words), dtype=torch.int32, device=dev.device)
    l_d, t_d, g_d = dev.alloc_views(n)
    t_d.fill_(1)
    m_d = torch.zeros((n, dev.n_roles), dtype=torch.int16, device=dev.device)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev.device)
    total = 0
    for step in range(steps):
        total += ref.playout_step(5, seed, step, s_r, l_r, t_r, g_r, m_r)
        dev.playout_step(s_d, l_d, t_d, g_d, seed, step, slot0=5, moves_out=m_d, counter=cnt)
        assert np.array_equal(_np(m_d), m_r), (game, step, 'moves')
        assert np.array_equal(_np(s_d), s_r), (game, step, 'states')
        assert np.array_equal(_np(l_d), l_r), (game, step, 'legal')
        assert np.array_equal(_np(t_d), t_r), (game, step, 'terminal')
        assert np.array_equal(_np(g_d), g_r), (game, step, 'goals')
    assert int(cnt.item()) == total


@pytest.mark.parametrize('game,n,steps', [('chess_15d', 100, 24), ('chess_15d', 1, 10), ('hexLG13', 40, 12)])
def test_wide_games_playout_bit_exact(game, n, steps):
    """BASELINE configs 3 and 5: games whose propnets exceed the 16-bit slot format (chess_15d: 243 K
    gates, hexLG13: 29 753 bases / 574 K gates) run through the wide kernel (32-bit literals, value
    array in L2-resident scratch) — same bit-exact contract against the C oracle."""
    dev = _sm(game)
    ref = RefStateMachine(tables.blob_by_name(game))
    assert dev.state_words == ref.state_words and dev.legal_words == ref.legal_words
    seed = 0xC0DE + n
    s_r = np.zeros((n, ref.state_words), dtype=np.uint32)
    l_r = np.zeros((n, ref.legal_words), dtype=np.uint32)
    t_r = np.ones(n, dtype=np.uint8)
    g_r = np.zeros((n, ref.n_roles), dtype=np.uint8)
    m_r = np.zeros((n, ref.n_roles), dtype=np.int16)
    s_d = torch.zeros((n, dev.state_words), dtype=torch.int32, device=dev.device)
    l_d, t_d, g_d = dev.alloc_views(n)
    t_d.fill_(1)
    m_d = torch.zeros((n, dev.n_roles), dtype=torch.int16, device=dev.device)
    prev = None
    for step in range(steps):
        prev = s_r.copy()
        ref.playout_step(0, seed, step, s_r, l_r, t_r, g_r, m_r)
        dev.playout_step(s_d, l_d, t_d, g_d, seed, step, moves_out=m_d)
        assert np.array_equal(_np(m_d), m_r), (game, step, 'moves')
        assert np.array_equal(_np(s_d), s_r), (game, step, 'states')
        assert np.array_equal(_np(l_d), l_r), (game, step, 'legal')
        assert np.array_equal(_np(t_d), t_r), (game, step, 'terminal')
        assert np.array_equal(_np(g_d), g_r), (game, step, 'goals')
    # transition + eval entry points on the last ply
    live = m_r[:, 0] >= 0
    got = dev.transition(torch.from_numpy(prev[live].view(np.int32)).to(dev.device),
                         torch.from_numpy(m_r[live]).to(dev.device))
    for a, b in zip(got, ref.transition_batch(prev[live], m_r[live])):
        assert np.array_equal(_np(a), b)
    for a, b in zip(dev.eval(s_d), ref.eval_batch(s_r)):
        assert np.array_equal(_np(a), b)


@pytest.mark.parametrize('game,n', [('breakthrough', 1), ('breakthrough', 33), ('breakthrough', 1000),
                                    ('breakthrough', 1024), ('breakthrough', 5000),
                                    ('breakthroughSmall', 33), ('breakthroughSmall', 3000)])
def test_specialised_kernel_bit_exact(game, n):
    """The generated thread-sliced kernels (csrc/gz_spec_<game>.cu), forced on for small
    batches: playout / transition / eval bit-exact with the C oracle, ragged tails included."""
    dev = _sm(game)
    assert dev.specialised, 'device program hash does not match the generated kernel: rerun tools/gen_spec_kernel.py'
    dev.set_spec_min_batch(0)
    ref = RefStateMachine(tables.blob_by_name(game))
    seed = 0x5BEC + n
    s_r = np.zeros((n, ref.state_words), dtype=np.uint32)
    l_r = np.zeros((n, ref.legal_words), dtype=np.uint32)
    t_r = np.ones(n, dtype=np.uint8)
    g_r = np.zeros((n, ref.n_roles), dtype=np.uint8)
    m_r = np.zeros((n, ref.n_roles), dtype=np.int16)
    s_d = torch.zeros((n, dev.state_words), dtype=torch.int32, device=dev.device)
    l_d, t_d, g_d = dev.alloc_views(n)
    t_d.fill_(1)
    m_d = torch.zeros((n, dev.n_roles), dtype=torch.int16, device=dev.device)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev.device)
    total = 0
    prev = None
    for step in range(80):
        prev = s_r.copy()
        total += ref.playout_step(3, seed, step, s_r, l_r, t_r, g_r, m_r)
        dev.playout_step(s_d, l_d, t_d, g_d, seed, step, slot0=3, moves_out=m_d, counter=cnt)
        assert np.array_equal(_np(m_d), m_r), (step, 'moves')
        assert np.array_equal(_np(s_d), s_r), (step, 'states')
        assert np.array_equal(_np(l_d), l_r), (step, 'legal')
        assert np.array_equal(_np(t_d), t_r), (step, 'terminal')
        assert np.array_equal(_np(g_d), g_r), (step, 'goals')
    assert int(cnt.item()) == total
    mv = m_r.copy()
    want = ref.transition_batch(prev, np.where(mv < 0, 0, mv).astype(np.int16))
    mv_dev = torch.from_numpy(mv).to(dev.device)
    got = dev.transition(torch.from_numpy(prev.view(np.int32)).to(dev.device), mv_dev)
    live = mv[:, 0] >= 0
    for a, b in zip(got, want):
        assert np.array_equal(_np(a)[live], b[live])
    # rows with a negative first move keep their state (the PUCT pipeline's root evaluation)
    assert np.array_equal(_np(got[0])[~live], prev[~live])
    for a, b in zip(dev.eval(s_d), ref.eval_batch(s_r)):
        assert np.array_equal(_np(a), b)
    # the table-driven kernel gives the same answers
    dev.set_spec_min_batch(-1)
    for a, b in zip(dev.eval(s_d), ref.eval_batch(s_r)):
        assert np.array_equal(_np(a), b)


@pytest.mark.parametrize('game,n,steps', [('breakthrough', 1100, 60), ('breakthroughSmall', 2048, 40), ('reversi', 1500, 66),
                                          ('connect6', 1024, 20), ('amazons_10x10', 1030, 24), ('chess_15d', 1100, 24),
                                          ('hexLG13', 1024, 8), ('reversi', 5, 20), ('chess_15d', 37, 12)])
def test_group_kernel_bit_exact(game, n, steps):
    """The group kernel (gz_propnet_group.cuh: 1024 games per CTA or cluster, one tile per lane) on full and
    ragged groups, forced also for tiny batches: playout plies (legal masks, moves, next states, terminal, goals,
    transition count) against the C oracle; then transition / eval entry points and the terminal-only goals mode
    against the per-tile kernels on the same positions."""
    dev = _sm(game)
    dev.set_group_min_batch(0)
    dev.set_spec_min_batch(-1)
    assert dev.launch_info is not None
    ref = RefStateMachine(tables.blob_by_name(game))
    seed = 0x6709 + n
    s_r = np.zeros((n, ref.state_words), dtype=np.uint32)
    l_r = np.zeros((n, ref.legal_words), dtype=np.uint32)
    t_r = np.ones(n, dtype=np.uint8)
    g_r = np.zeros((n, ref.n_roles), dtype=np.uint8)
    m_r = np.zeros((n, ref.n_roles), dtype=np.int16)
    s_d = torch.zeros((n, dev.state_words), dtype=torch.int32, device=dev.device)
    l_d, t_d, g_d = dev.alloc_views(n)
    t_d.fill_(1)
    m_d = torch.zeros((n, dev.n_roles), dtype=torch.int16, device=dev.device)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev.device)
    total = 0
    for step in range(steps):
        prev_d = s_d.clone()
        total += ref.playout_step(3, seed, step, s_r, l_r, t_r, g_r, m_r)
        dev.playout_step(s_d, l_d, t_d, g_d, seed, step, slot0=3, moves_out=m_d, counter=cnt)
        assert np.array_equal(_np(m_d), m_r), (game, step, 'moves')
        assert np.array_equal(_np(s_d), s_r), (game, step, 'states')
        assert np.array_equal(_np(l_d), l_r), (game, step, 'legal')
        assert np.array_equal(_np(t_d), t_r), (game, step, 'terminal')
        assert np.array_equal(_np(g_d), g_r), (game, step, 'goals')
    assert int(cnt.item()) == total
    # transition with given moves (rows restarted by the last ply: identity move -1) and eval: group vs tile kernels
    mv = m_d.clone()
    out_g = dev.transition(prev_d, mv)
    ev_g = dev.eval(s_d)
    dev.set_group_min_batch(-1)
    out_t = dev.transition(prev_d, mv)
    ev_t = dev.eval(s_d)
    for a, b in zip(out_g + ev_g, out_t + ev_t):
        assert torch.equal(a, b)
    live = (m_r[:, 0] >= 0)
    assert np.array_equal(_np(out_g[0])[live], s_r[live])
    # terminal-only goals
    dev2 = _sm(game, goals_mode=gzlib.GOALS_TERMINAL_ONLY)
    dev2.set_group_min_batch(0)
    dev2.set_spec_min_batch(-1)
    lg, tm, gl = dev2.eval(s_d)
    gl = _np(gl)
    assert np.array_equal(_np(tm), t_r) and np.array_equal(_np(lg), l_r)
    assert np.array_equal(gl[t_r == 1], g_r[t_r == 1]) and (gl[t_r == 0] == 255).all()


@pytest.mark.parametrize('game', ['breakthrough', 'reversi'])
def test_transition_eval_and_host_api(game):
    dev = _sm(game)
    ref = RefStateMachine(tables.blob_by_name(game))
    n = 777
    s = np.zeros((n, ref.state_words), dtype=np.uint32)
    l = np.zeros((n, ref.legal_words), dtype=np.uint32)
    t = np.ones(n, dtype=np.uint8)
    g = np.zeros((n, ref.n_roles), dtype=np.uint8)
    m = np.zeros((n, ref.n_roles), dtype=np.int16)
    for step in range(25):
        prev = s.copy()
        ref.playout_step(0, 99, step, s, l, t, g, m)
    # last step: prev --m--> s for the slots that were not restarted
    live = (m[:, 0] >= 0)
    assert live.sum() > n // 2
    prev_l, mv = prev[live], m[live]
    want = ref.transition_batch(prev_l, mv)
    got = dev.transition(torch.from_numpy(prev_l.view(np.int32)).to(dev.device),
                         torch.from_numpy(mv).to(dev.device))
    for a, b in zip(got, want):
        assert np.array_equal(_np(a), b)
    got_h = dev.transition_host(prev_l, mv)
    for a, b in zip(got_h, want):
        assert np.array_equal(a, b)
    ev = dev.eval(torch.from_numpy(s.view(np.int32)).to(dev.device))
    for a, b in zip(ev, ref.eval_batch(s)):
        assert np.array_equal(_np(a), b)


def test_goals_terminal_only_mode():
    game = 'breakthroughSmall'
    dev = _sm(game, goals_mode=gzlib.GOALS_TERMINAL_ONLY)
    ref = RefStateMachine(tables.blob_by_name(game))
    n = 256
    s = np.zeros((n, ref.state_words), dtype=np.uint32)
    l = np.zeros((n, ref.legal_words), dtype=np.uint32)
    t = np.ones(n, dtype=np.uint8)
    g = np.zeros((n, ref.n_roles), dtype=np.uint8)
    for step in range(45):
        ref.playout_step(0, 3, step, s, l, t, g, None)
    assert t.any() and not t.all()
    lg, tm, gl = dev.eval(torch.from_numpy(s.view(np.int32)).to(dev.device))
    gl = _np(gl)
    assert np.array_equal(_np(tm), t)
    assert np.array_equal(gl[t == 1], g[t == 1])
    assert (gl[t == 0] == 255).all()


def test_deferred_goal_pass_bit_exact():
    """reversi, goals for terminal states only: the compacted second pass over the terminal states (default from 4096
    states up) returns what the one-pass kernel and the C oracle return — goal values of every terminal state, 255
    elsewhere — on playout plies and on transitions with given moves."""
    game = 'reversi'
    a = _sm(game, goals_mode=gzlib.GOALS_TERMINAL_ONLY)
    b = _sm(game, goals_mode=gzlib.GOALS_TERMINAL_ONLY)
    b.set_deferred_goals_min_batch(-1)
    ref = RefStateMachine(tables.blob_by_name(game))
    n = 8192 + 37
    sa = torch.zeros((n, a.state_words), dtype=torch.int32, device=a.device)
    sb = torch.zeros_like(sa)
    la, ta, ga = a.alloc_views(n)
    lb, tb, gb = b.alloc_views(n)
    ta.fill_(1); tb.fill_(1)
    s = np.zeros((n, ref.state_words), dtype=np.uint32)
    l = np.zeros((n, ref.legal_words), dtype=np.uint32)
    t = np.ones(n, dtype=np.uint8)
    g = np.zeros((n, ref.n_roles), dtype=np.uint8)
    seen_terminal = 0
    for step in range(75):
        a.playout_step(sa, la, ta, ga, 11, step)
        b.playout_step(sb, lb, tb, gb, 11, step)
        ref.playout_step(0, 11, step, s, l, t, g, None)
        if step % 6 == 5 or step > 55:
            assert np.array_equal(_np(sa).view(np.uint32), s) and np.array_equal(_np(ta), t)
            assert np.array_equal(_np(ga), _np(gb))
            gl = _np(ga)
            assert np.array_equal(gl[t == 1], g[t == 1]) and (gl[t == 0] == 255).all()
            seen_terminal += int(t.sum())
    assert seen_terminal > 100
    # transition with given moves from the last positions (terminal rows: identity move)
    mv = torch.full((n, a.n_roles), -1, dtype=torch.int16, device=a.device)
    na, la2, ta2, ga2 = a.transition(sa, mv)
    nb, lb2, tb2, gb2 = b.transition(sb, mv)
    assert np.array_equal(_np(ga2), _np(gb2)) and np.array_equal(_np(ta2), _np(tb2))
    assert (_np(ga2)[_np(ta2) == 1] != 255).any()


def test_empty_and_bad_arguments():
    dev = _sm('breakthrough')
    s = torch.zeros((0, dev.state_words), dtype=torch.int32, device=dev.device)
    lg, tm, gl = dev.eval(s)
    assert lg.shape[0] == 0
    L = gzlib.lib()
    rc = L.gz_sm_eval(dev._h, 4, None, None, None, None, 0, None)
    assert rc != 0 and b'null' in L.gz_last_error()


def test_full_size_playout_properties():
    """BASELINE config 1 size (2^20 slots): size-independent checks — transition count equals the
    number of non-terminal slots, restarted slots hold the initial state, legal masks non-empty."""
    dev = _sm('breakthrough')
    n = 1 << 20
    s = torch.zeros((n, dev.state_words), dtype=torch.int32, device=dev.device)
    l, t, g = dev.alloc_views(n)
    t.fill_(1)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev.device)
    expect = 0
    for step in range(70):
        live = int((t == 0).sum().item())
        expect += live
        was_term = t.clone()
        dev.playout_step(s, l, t, g, 0xB200, step, counter=cnt)
        if step in (0, 69):
            init = torch.from_numpy(dev.init_words.view(np.int32)).to(dev.device)
            restarted = s[was_term == 1]
            assert (restarted == init).all()
    assert int(cnt.item()) == expect
    # every non-terminal state has at least one legal move per role
    lw = l[t == 0]
    for r in range(dev.n_roles):
        words = lw[:, dev.legal_word_off[r]:dev.legal_word_off[r + 1]]
        assert (words != 0).any(dim=1).all()
    # bit-exact against the oracle on a 2^12-slot slice of the same run (same slot ids and seed)
    ref = RefStateMachine(tables.blob_by_name('breakthrough'))
    m = 1 << 12
    sr = np.zeros((m, ref.state_words), dtype=np.uint32)
    lr = np.zeros((m, ref.legal_words), dtype=np.uint32)
    tr = np.ones(m, dtype=np.uint8)
    gr = np.zeros((m, ref.n_roles), dtype=np.uint8)
    off = 123 * 32
    for step in range(70):
        ref.playout_step(off, 0xB200, step, sr, lr, tr, gr, None)
    assert np.array_equal(_np(s[off:off + m]), sr)
    assert np.array_equal(_np(l[off:off + m]), lr)
    assert np.array_equal(_np(t[off:off + m]), tr)
