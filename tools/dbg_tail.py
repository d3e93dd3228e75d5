"""Which game warps set the duration of puct_pre_kernel? Per-warp cycle counts (gz_puct_set_debug_cycles)."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch
from galvanise_zero_b200.engine import PuctEngine
from galvanise_zero_b200.defs import confs
G = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
GAME = sys.argv[2] if len(sys.argv) > 2 else 'reversi'
CC = int(sys.argv[4]) if len(sys.argv) > 4 else 4096 * 16
sp = confs.selfplay_config_template(evals_per_move=800)
eng = PuctEngine(GAME, sp, n_games=G, mode=0, node_cap=4096, child_cap=CC)
eng.run_synthetic(int(sys.argv[1]) if len(sys.argv) > 1 else 3000)
buf = torch.zeros((G, 4), dtype=torch.int64, device=eng.device)
eng._L.gz_puct_set_debug_cycles.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
eng._L.gz_puct_set_debug_cycles(eng._h, buf.data_ptr())
rows = []
for it in range(40):
    eng.run_synthetic(1)
    torch.cuda.synchronize()
    rows.append(buf.cpu().numpy().copy())
eng._L.gz_puct_set_debug_cycles(eng._h, None)
a = np.stack(rows)            # [40, G, 4]
cyc = a[:, :, 0].astype(np.float64) / 1.965e3   # us
print('per-step max us:', np.round(cyc.max(axis=1), 1))
print('per-step mean us: %.1f  median %.1f  p90 %.1f  p99 %.1f  max %.1f' % (cyc.mean(), np.median(cyc), np.percentile(cyc, 90), np.percentile(cyc, 99), cyc.max()))
# what do the slow warps do?
flat = a.reshape(-1, 4)
c = flat[:, 0] / 1.965e3
order = np.argsort(-c)[:40]
for i in order:
    ph = int(flat[i, 1]); print('us %.1f  phase %d -> %d  playouts %d  moves %d  path_len %d  req %d' % (c[i], ph // 256, ph % 256, flat[i, 2], flat[i, 3] // 65536, (flat[i, 3] // 256) % 256, flat[i, 3] % 256))
for lo, hi in [(0, 20), (20, 40), (40, 80), (80, 160), (160, 1e9)]:
    m = (c >= lo) & (c < hi)
    if m.sum():
        print('bucket %4d-%4s us: %6.2f%% of warps, mean playouts %.2f, moved %.3f' % (lo, hi if hi < 1e9 else 'inf', 100 * m.mean(), flat[m, 2].mean(), (flat[m, 3] // 65536).mean()))
