import sys; sys.path.insert(0,'/root/repo')
import torch
from galvanise_zero_b200.engine import PuctEngine
from galvanise_zero_b200.defs import confs
conf = confs.base_puct_config(dirichlet_noise_pct=-1, choose='choose_top_visits')
sp = confs.selfplay_config_template(); sp.puct_config = conf
eng = PuctEngine('reversi', sp, n_games=1, mode=1, node_cap=8192, child_cap=1<<18, plane_dtype=torch.float32)
eng.command(0, 1)
for i in range(6):
    eng.run_synthetic(1)
    s = eng.snapshot(0); t = eng.totals()
    print(i, {k: s[k] for k in ('cmd_done','phase','error','nodes','children','root_visits')}, 'waiting', t['waiting_eval'], 'evals', t['evals'])
eng.command(0, 3, evals=50)
for i in range(8):
    eng.run_synthetic(1)
    s = eng.snapshot(0); t = eng.totals()
    print(i, {k: s[k] for k in ('cmd_done','phase','error','nodes','children','root_visits')}, 'waiting', t['waiting_eval'], 'evals', t['evals'], 'playouts', t['playouts'])
