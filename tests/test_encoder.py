"""Encoder parity. CPU: the C oracle (restatement of gdltransformer.cpp) agrees with the Python
reference encoder restated from nn/bases.py:242-287 on golden-fixture states, and the plane shapes
match the reference's shipped models. GPU: gz_encode is bit-exact with the oracle (f32 and bf16)."""
import numpy as np
import pytest

from galvanise_zero_b200 import bases, cppinterface, lookup, tables
from galvanise_zero_b200.defs import datadesc
from oracle.pnref import RefEncoder, RefStateMachine

GAMES = ['breakthrough', 'reversi', 'connect6', 'amazons_10x10', 'breakthroughSmall', 'chess_15d', 'hexLG13']
SHAPES = {'breakthrough': (5, 8, 8), 'reversi': (5, 8, 8), 'connect6': (8, 19, 19),
          'amazons_10x10': (12, 10, 10), 'chess_15d': (32, 8, 8),
          'hexLG13': (5, 13, 13)}          # data/*/models/*.json batch_input_shape of the reference


def _setup(game, num_prev=1):
    gi = lookup.by_name(game)
    tr = bases.GdlBasesTransformer(gi, datadesc.generation_for(game, num_previous_states=num_prev))
    ct = cppinterface.create_c_transformer(tr)
    return gi, tr, ct


def _trajectory(game, n_games=None, plies=None, seed=11):
    big = game in ('chess_15d', 'hexLG13')
    n_games = n_games or (3 if big else 6)
    plies = plies or (16 if big else 40)
    ref = RefStateMachine(tables.blob_by_name(game))
    s = np.zeros((n_games, ref.state_words), dtype=np.uint32)
    l = np.zeros((n_games, ref.legal_words), dtype=np.uint32)
    t = np.ones(n_games, dtype=np.uint8)
    g = np.zeros((n_games, ref.n_roles), dtype=np.uint8)
    hist = []
    for step in range(plies):
        ref.playout_step(0, seed, step, s, l, t, g, None)
        hist.append(s.copy())
    return hist


@pytest.mark.parametrize('game', GAMES)
def test_oracle_encoder_matches_python_reference_encoder(game):
    gi, tr, ct = _setup(game)
    if game in SHAPES:
        assert (tr.num_channels, tr.num_cols, tr.num_rows) == SHAPES[game]
    assert ct.total_size() == tr.num_channels * tr.channel_size
    enc = RefEncoder.from_c_transformer(ct)
    hist = _trajectory(game)
    nb = tr.num_bases
    for k in range(0, len(hist), 3):
        cur = hist[k]
        prev = hist[k - 1] if k > 0 else None
        if prev is not None:
            got = enc.encode(cur, prev[:, None, :])
        else:
            got = enc.encode(cur)
        for i in range(cur.shape[0]):
            bits = tables.unpack_state_bits(cur[i], nb)
            pbits = [tables.unpack_state_bits(prev[i], nb)] if prev is not None else []
            want = tr.state_to_channels(bits, pbits).reshape(-1)
            assert np.array_equal(got[i], want)


def test_oracle_prev_state_quirk_all_prev_land_in_group_one():
    """gdltransformer.cpp:39-43: `count` is never incremented."""
    gi, tr, ct = _setup('breakthroughSmall', num_prev=2)
    enc = RefEncoder.from_c_transformer(ct)
    hist = _trajectory('breakthroughSmall', n_games=3, plies=10)
    cur, p1, p2 = hist[9], hist[8], hist[7]
    prev = np.stack([p1, p2], axis=1)
    got = enc.encode(cur, prev)
    gs = tr.raw_channels_per_state * tr.channel_size
    one = enc.encode(p1)[:, :gs]
    two = enc.encode(p2)[:, :gs]
    assert np.array_equal(got[:, gs:2 * gs], np.maximum(one, two))
    assert not got[:, 2 * gs:3 * gs].any()


@pytest.mark.parametrize('game', GAMES)
def test_oracle_encoder_is_the_references_compiled_to_channels(game):
    """oracle/encoder_ref.c against the reference's own GdlBasesTransformer::toChannels (gdltransformer.cpp:22-52,
    compiled unmodified into oracle/_ref): same float32 planes on random-played positions with 0, 1 (and for
    one game 2) previous states — this pins the restatement the GPU encoder is checked against."""
    from oracle import gzref
    if not gzref.available():
        pytest.skip('oracle/_ref cannot be built here')
    for num_prev in ((1, 2) if game == 'breakthroughSmall' else (1,)):
        gi, tr, ct = _setup(game, num_prev=num_prev) if num_prev != 1 else _setup(game)
        enc = RefEncoder.from_c_transformer(ct)
        ref = gzref.RefEngine(tables.blob_by_name(game), ct)
        hist = _trajectory(game, n_games=23, plies=16)
        for k in (0, 3, 15):
            cur = hist[k]
            prev = cnt = None
            if k >= num_prev:
                prev = np.stack([hist[k - 1 - j] for j in range(num_prev)], axis=1)
                cnt = (np.arange(cur.shape[0]) % (num_prev + 1)).astype(np.uint8)
            assert np.array_equal(enc.encode(cur, prev, cnt), ref.encode(cur, prev, cnt))


@pytest.mark.gpu
@pytest.mark.parametrize('game', GAMES)
def test_gpu_encoder_is_the_references_compiled_to_channels(game):
    """gz_encode on the device against the reference's compiled toChannels directly (f32 exact, bf16 = rounded f32)."""
    import torch
    from oracle import gzref
    if not gzref.available():
        pytest.skip('oracle/_ref missing')
    gi, tr, ct = _setup(game)
    ct.finalize(tr.num_bases, 0)
    ref = gzref.RefEngine(tables.blob_by_name(game), ct)
    hist = _trajectory(game, n_games=29, plies=20)
    dev = torch.device('cuda', 0)
    for k in (0, 7, 19):
        cur = hist[k]
        prev = hist[k - 1][:, None, :] if k > 0 else None
        cnt = (np.arange(cur.shape[0]) % 2).astype(np.uint8) if k > 0 else None
        want = ref.encode(cur, prev, cnt)
        d_prev = torch.from_numpy(prev.view(np.int32).copy()).to(dev) if prev is not None else None
        d_cnt = torch.from_numpy(cnt).to(dev) if cnt is not None else None
        got = ct.encode(torch.from_numpy(cur.view(np.int32)).to(dev), d_prev, d_cnt, dtype=torch.float32).cpu().numpy()
        assert np.array_equal(got, want)
        got16 = ct.encode(torch.from_numpy(cur.view(np.int32)).to(dev), d_prev, d_cnt, dtype=torch.bfloat16)
        assert torch.equal(got16.cpu(), torch.from_numpy(want).to(torch.bfloat16))


@pytest.mark.gpu
@pytest.mark.parametrize('game', GAMES)
def test_gpu_encoder_bit_exact(game):
    import torch
    gi, tr, ct = _setup(game)
    ct.finalize(tr.num_bases, 0)
    enc = RefEncoder.from_c_transformer(ct)
    hist = _trajectory(game, n_games=37, plies=30)
    dev = torch.device('cuda', 0)
    for k in (0, 1, 14, 29):
        cur = hist[k]
        prev = hist[k - 1][:, None, :] if k > 0 else None
        cnt = None
        if prev is not None:
            cnt = (np.arange(cur.shape[0]) % 2).astype(np.uint8)     # some roots have no parent
        want = enc.encode(cur, prev, cnt)
        d_cur = torch.from_numpy(cur.view(np.int32)).to(dev)
        d_prev = torch.from_numpy(prev.view(np.int32).copy()).to(dev) if prev is not None else None
        d_cnt = torch.from_numpy(cnt).to(dev) if cnt is not None else None
        got32 = ct.encode(d_cur, d_prev, d_cnt, dtype=torch.float32).cpu().numpy()
        assert np.array_equal(got32, want)
        got16 = ct.encode(d_cur, d_prev, d_cnt, dtype=torch.bfloat16)
        want16 = torch.from_numpy(want).to(torch.bfloat16)
        assert torch.equal(got16.cpu(), want16)


@pytest.mark.gpu
def test_gpu_encoder_multi_prev_quirk_and_tail():
    import torch
    gi, tr, ct = _setup('breakthroughSmall', num_prev=2)
    ct.finalize(tr.num_bases, 0)
    enc = RefEncoder.from_c_transformer(ct)
    hist = _trajectory('breakthroughSmall', n_games=13, plies=12)
    cur = hist[11]
    prev = np.stack([hist[10], hist[9]], axis=1)
    cnt = (np.arange(13) % 3).astype(np.uint8)
    want = enc.encode(cur, prev, cnt)
    dev = torch.device('cuda', 0)
    got = ct.encode(torch.from_numpy(cur.view(np.int32)).to(dev),
                    torch.from_numpy(prev.view(np.int32).copy()).to(dev),
                    torch.from_numpy(cnt).to(dev), dtype=torch.float32).cpu().numpy()
    assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize('game,c_pad', [('reversi', 8), ('reversi', 16), ('amazons_10x10', 16)])
def test_gpu_encoder_channels_last_padded_layout(game, c_pad):
    """The additive [n][H][W][c_pad] layout holds exactly the reference planes, permuted, plus zero planes."""
    import torch
    gi, tr, ct = _setup(game)
    ct.finalize(tr.num_bases, 0)
    enc = RefEncoder.from_c_transformer(ct)
    hist = _trajectory(game, n_games=19, plies=12)
    cur, prev = hist[11], hist[10][:, None, :]
    want = enc.encode(cur, prev).reshape(19, tr.num_channels, tr.channel_size)
    dev = torch.device('cuda', 0)
    got = ct.encode(torch.from_numpy(cur.view(np.int32)).to(dev), torch.from_numpy(prev.view(np.int32).copy()).to(dev),
                    dtype=torch.bfloat16, c_pad=c_pad).float().cpu().numpy().reshape(19, tr.channel_size, c_pad)
    assert np.array_equal(got[:, :, :tr.num_channels], want.transpose(0, 2, 1))
    assert not got[:, :, tr.num_channels:].any()


# ------------------------------------------------------------------------------------------------
# encode_col_kernel (csrc/gz_encoder_col.cuh): the column-owned kernel for large batches
COL_GAMES = ['reversi', 'breakthrough', 'connect6', 'amazons_10x10', 'chess_15d']      # rows of a multiple of 8 elements
RAGGED_GAMES = ['breakthroughSmall', 'hexLG13']      # rows of 180 / 845 elements: units of 2 / 8 states


def _ctasim_encode(ct, n_bases, cur, prev, cnt, dtype, c_pad=0, S=8, grid=2):
    """The kernel's own source compiled for the CPU (tests/hostsim/ctasim_encoder.cpp: 256 threads per CTA)."""
    import ctypes
    import os
    import subprocess
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'hostsim')
    subprocess.check_call(['make', '-s', '-C', here, 'libctasim.so'])
    L = ctypes.CDLL(os.path.join(here, 'libctasim.so'))
    vp, ci = ctypes.c_void_p, ctypes.c_int
    L.cs_encode_col.argtypes = [vp, ci, vp, vp, vp, ci, ci, ci, ci, ctypes.c_int64, ci, vp, vp, vp, ci, vp, ci, ci, ci, ci]
    group = ct.channels_per_state * ct.channel_size
    src = np.full(group, -1, dtype=np.int32)
    for b, off in ct.board_bases:                      # (gz_transformer_finalize builds the same table)
        src[off] = b
    cb = np.array([b for b, c, v in ct.control_bases], dtype=np.int32)
    cc = np.array([c for b, c, v in ct.control_bases], dtype=np.int32)
    cv = np.array([v for b, c, v in ct.control_bases], dtype=np.float32)
    cur = np.ascontiguousarray(cur, dtype=np.uint32)
    n, sw = cur.shape
    total = ct.total_size()
    row = ct.channel_size * c_pad if c_pad else total
    out = np.full((n, row), 0x7777 if dtype == 1 else 0x77777777, dtype=np.uint16 if dtype == 1 else np.uint32)
    max_prev = 0
    if prev is not None:
        prev = np.ascontiguousarray(prev, dtype=np.uint32)
        max_prev = prev.shape[1]
    rc = L.cs_encode_col(src.ctypes.data, group, cb.ctypes.data, cc.ctypes.data, cv.ctypes.data, len(cb),
                         ct.channel_size, ct.num_prev_states, ct.num_control_channels, n, sw, cur.ctypes.data,
                         prev.ctypes.data if prev is not None else None,
                         cnt.ctypes.data if cnt is not None else None, max_prev, out.ctypes.data, dtype, c_pad, S, grid)
    assert rc == 0
    return out


def _bf16_bits(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)


@pytest.mark.timeout(600)
@pytest.mark.parametrize('game', COL_GAMES + RAGGED_GAMES)
def test_column_kernel_source_under_cta_emulation_is_bit_exact(game):
    gi, tr, ct = _setup(game)
    enc = RefEncoder.from_c_transformer(ct)
    big = game in ('connect6', 'chess_15d', 'hexLG13')
    n = 13 if big else 37                      # (ragged rows: 13 = one unit of 8 + 5 states written element-wise)
    hist = _trajectory(game, n_games=n, plies=9)
    cur, prev = hist[8], hist[7][:, None, :]
    cnt = (np.arange(n) % 2).astype(np.uint8)
    want = enc.encode(cur, prev, cnt)
    got = _ctasim_encode(ct, tr.num_bases, cur, prev, cnt, 0, S=8, grid=2)          # ragged last tile, grid stride
    assert np.array_equal(got.view(np.float32), want)
    got16 = _ctasim_encode(ct, tr.num_bases, cur, prev, cnt, 1, S=5 if game in COL_GAMES else 16, grid=1)
    assert np.array_equal(got16, _bf16_bits(want))
    got16 = _ctasim_encode(ct, tr.num_bases, cur, prev, cnt, 1, S=64, grid=1)        # one tile holds the batch
    assert np.array_equal(got16, _bf16_bits(want))
    want0 = enc.encode(cur)                                                          # no previous states at all
    assert np.array_equal(_ctasim_encode(ct, tr.num_bases, cur, None, None, 0, S=16, grid=3).view(np.float32), want0)
    if not big and game in COL_GAMES:
        c_pad = 16
        w3 = enc.encode(cur, prev).reshape(n, tr.num_channels, tr.channel_size)
        g3 = _ctasim_encode(ct, tr.num_bases, cur, prev, None, 1, c_pad=c_pad, S=8, grid=2).reshape(n, tr.channel_size, c_pad)
        assert np.array_equal(g3[:, :, :tr.num_channels], _bf16_bits(w3.transpose(0, 2, 1)))
        assert not g3[:, :, tr.num_channels:].any()


@pytest.mark.gpu
@pytest.mark.parametrize('game', COL_GAMES + RAGGED_GAMES)
def test_gpu_column_kernel_bit_exact(game):
    """gz_encode with the column kernel forced (any batch size) against the C oracle: f32 exact, bf16 = rounded
    f32, ragged batch, some roots without a previous state, the padded channels-last layout; and the default
    dispatch (a batch above the threshold) gives the same planes as the small-batch kernels."""
    import ctypes
    import torch
    from galvanise_zero_b200 import lib as gzlib
    gi, tr, ct = _setup(game)
    ct.finalize(tr.num_bases, 0)
    enc = RefEncoder.from_c_transformer(ct)
    L = gzlib.lib()
    L.gz_encode_set_column_kernel_min_batch.argtypes = [ctypes.c_int64]
    L.gz_encode_set_column_kernel_ragged(1 if game in RAGGED_GAMES else 0)
    n = 331 if game != 'hexLG13' else 45
    hist = _trajectory(game, n_games=n, plies=12)
    dev = torch.device('cuda', 0)
    try:
        L.gz_encode_set_column_kernel_min_batch(1)
        for k, tile in zip((0, 5, 11, 11), (0, 16, 0, 64)):
            L.gz_encode_set_column_kernel_tile(tile)
            cur = hist[k]
            prev = hist[k - 1][:, None, :] if k > 0 else None
            cnt = (np.arange(n) % 3 != 0).astype(np.uint8) if k > 0 else None
            want = enc.encode(cur, prev, cnt)
            d_cur = torch.from_numpy(cur.view(np.int32)).to(dev)
            d_prev = torch.from_numpy(prev.view(np.int32).copy()).to(dev) if prev is not None else None
            d_cnt = torch.from_numpy(cnt).to(dev) if cnt is not None else None
            assert np.array_equal(ct.encode(d_cur, d_prev, d_cnt, dtype=torch.float32).cpu().numpy(), want)
            assert torch.equal(ct.encode(d_cur, d_prev, d_cnt, dtype=torch.bfloat16).cpu(),
                               torch.from_numpy(want).to(torch.bfloat16))
            if k == 11 and game in COL_GAMES:
                cp = 16 if tr.num_channels <= 16 else 32
                w3 = want.reshape(n, tr.num_channels, tr.channel_size)
                g3 = ct.encode(d_cur, d_prev, d_cnt, dtype=torch.bfloat16, c_pad=cp).float().cpu().numpy()
                g3 = g3.reshape(n, tr.channel_size, cp)
                assert np.array_equal(g3[:, :, :tr.num_channels], w3.transpose(0, 2, 1)) and not g3[:, :, tr.num_channels:].any()
        # default dispatch: a batch above the threshold takes the column kernel, below it the small-batch kernels
        L.gz_encode_set_column_kernel_min_batch(16384)
        reps = 16384 // n + 1
        big = torch.from_numpy(np.tile(hist[11], (reps, 1)).view(np.int32)).to(dev)
        bigp = torch.from_numpy(np.tile(hist[10], (reps, 1)).view(np.int32).copy()).to(dev).reshape(-1, 1, big.shape[1])
        a = ct.encode(big, bigp, dtype=torch.bfloat16)
        L.gz_encode_set_column_kernel_min_batch(-1)
        b = ct.encode(big, bigp, dtype=torch.bfloat16)
        assert a.shape[0] >= 16384 and torch.equal(a, b)
    finally:
        L.gz_encode_set_column_kernel_min_batch(16384)
        L.gz_encode_set_column_kernel_tile(0)
        L.gz_encode_set_column_kernel_ragged(-1)
