"""Encoder throughput per game and column-kernel tuning (measurement aid): python tools/bench_encoder.py [games..]"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from galvanise_zero_b200 import bases, cppinterface, lookup, lib as gzlib      # noqa: E402
from galvanise_zero_b200.defs import datadesc                                   # noqa: E402
from galvanise_zero_b200.statemachine import DeviceStateMachine                 # noqa: E402

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs'] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')) else 6650.0
L = gzlib.lib()
L.gz_encode_set_column_kernel_tile.argtypes = [ctypes.c_int]
L.gz_encode_set_column_kernel_min_batch.argtypes = [ctypes.c_int64]
RAGGED = os.environ.get('GZ_BENCH_RAGGED', '') == '1'         # rows that are not a multiple of 8 elements through the column kernel
L.gz_encode_set_column_kernel_ragged(1 if RAGGED else -1)
SLOTS = int(os.environ.get('GZ_BENCH_SLOTS', str(1 << 19)))
for game in (sys.argv[1:] or ['reversi', 'connect6', 'amazons_10x10', 'chess_15d']):
    dev = DeviceStateMachine(game, device=0)
    n = SLOTS
    s = torch.zeros((n, dev.state_words), dtype=torch.int32, device=dev.device)
    l, t, g = dev.alloc_views(n)
    t.fill_(1)
    for step in range(12):
        dev.playout_step(s, l, t, g, 0xB200, step)
    tr = bases.GdlBasesTransformer(lookup.by_name(game), datadesc.generation_for(game))
    ct = cppinterface.create_c_transformer(tr, device=0)
    prev = s.clone().reshape(n, 1, dev.state_words)
    eb = 2 * ((dev.n_bases + 7) // 8) + 2 * ct.total_size()
    for dtype in (torch.bfloat16, torch.float32):
        out = torch.empty((n, ct.total_size()), dtype=dtype, device=dev.device)
        bytes_per = 2 * ((dev.n_bases + 7) // 8) + out.element_size() * ct.total_size()
        configs = (('encode_fast_kernel', -1, 0), ('encode_col_kernel S=32', 1, 32), ('encode_col_kernel S=64', 1, 64),
                   ('encode_col_kernel S=128', 1, 128), ('encode_col_kernel S=256', 1, 256), ('encode_col_kernel default', 1, 0))
        if RAGGED:
            configs = (('small-batch kernels', -1, 0), ('encode_col_kernel S=8', 1, 8), ('encode_col_kernel S=16', 1, 16),
                       ('encode_col_kernel default', 1, 0))
        for label, mb, tile in configs:
            L.gz_encode_set_column_kernel_min_batch(mb)
            L.gz_encode_set_column_kernel_tile(tile)
            for _ in range(3):
                ct.encode(s, prev, out=out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                ct.encode(s, prev, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            ach = bytes_per * n / (ms * 1e-3) / 1e9
            print(json.dumps({'game': game, 'dtype': str(dtype).split('.')[-1], 'kernel': label, 'ms': round(ms, 4),
                              'GBps': round(ach, 1), 'frac_hbm': round(ach / PEAK, 3), 'bytes_per_position': bytes_per}), flush=True)
    del s, l, t, g, prev, out
    torch.cuda.empty_cache()
L.gz_encode_set_column_kernel_min_batch(16384)
L.gz_encode_set_column_kernel_tile(0)
L.gz_encode_set_column_kernel_ragged(-1)
