"""Profiling helper: PUCT engine steps with the synthetic network (no PyTorch kernels in the way)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from galvanise_zero_b200.engine import PuctEngine
from galvanise_zero_b200.defs import confs
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
pre = int(sys.argv[2]) if len(sys.argv) > 2 else 600
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
sp = confs.selfplay_config_template(evals_per_move=int(sys.argv[4]) if len(sys.argv) > 4 else 800)
if len(sys.argv) > 5: sp.oscillate_sampling_pct = float(sys.argv[5])
eng = PuctEngine('reversi', sp, n_games=G, mode=0, node_cap=4096, child_cap=4096 * 16)
eng.run_synthetic(pre)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.run_synthetic(steps)
e1.record()
torch.cuda.synchronize()
print('ms/step', e0.elapsed_time(e1) / steps, eng.totals())
