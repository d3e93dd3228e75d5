"""One game's default encoder launches (for an ncu capture): python tools/bench_encoder_one.py reversi"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from galvanise_zero_b200 import bases, cppinterface, lookup      # noqa: E402
from galvanise_zero_b200.defs import datadesc                     # noqa: E402
from galvanise_zero_b200.statemachine import DeviceStateMachine   # noqa: E402

game = sys.argv[1] if len(sys.argv) > 1 else 'reversi'
dev = DeviceStateMachine(game, device=0)
n = 1 << 19
s = torch.zeros((n, dev.state_words), dtype=torch.int32, device=dev.device)
l, t, g = dev.alloc_views(n)
t.fill_(1)
for step in range(12):
    dev.playout_step(s, l, t, g, 0xB200, step)
tr = bases.GdlBasesTransformer(lookup.by_name(game), datadesc.generation_for(game))
ct = cppinterface.create_c_transformer(tr, device=0)
prev = s.clone().reshape(n, 1, dev.state_words)
out = torch.empty((n, ct.total_size()), dtype=torch.bfloat16, device=dev.device)
for _ in range(6):
    ct.encode(s, prev, out=out)
torch.cuda.synchronize()
print('ok', game, n)
