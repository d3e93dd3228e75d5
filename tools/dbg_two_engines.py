"""Two engines on two streams with the synthetic network (no graphs): concurrency sanity check."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from galvanise_zero_b200.engine import PuctEngine
from galvanise_zero_b200.defs import confs
G = int(sys.argv[1]); steps = int(sys.argv[2])
sp = confs.selfplay_config_template(evals_per_move=800)
engs = [PuctEngine('reversi', sp, n_games=G, mode=0, node_cap=4096, child_cap=4096 * 16, seed=0xB200 + 104729 * k) for k in range(2)]
streams = [torch.cuda.Stream() for _ in engs]
t0 = time.time()
for i in range(steps):
    for e, s in zip(engs, streams):
        with torch.cuda.stream(s):
            e.run_synthetic(1)
    if i % 500 == 499:
        torch.cuda.synchronize()
        print(i, time.time() - t0, [e.totals()['moves'] for e in engs], flush=True)
torch.cuda.synchronize()
print('done', time.time() - t0, [e.totals() for e in engs][0])
