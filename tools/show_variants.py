"""Prints one line per bench JSON: throughput, regime and per-kernel times (tuning aid)."""
import glob
import json
import sys
for f in sorted(glob.glob(sys.argv[1])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'ERR', e)
        continue
    k = d['kernel_times_ms']
    r = d['regime']
    print('%-34s moves/s %6.0f evals/s %5.2fM ms/adv %.3f fill %.2f po/ev %.2f | pre %.3f rr %.3f pn %.3f post %.3f enc %.3f nn %.3f' % (
        f.split('/')[-1], d['value'], d['evals_per_s'] / 1e6, d['config']['ms_per_advance'], r['nn_batch_fill'],
        r['playouts_per_eval'], k['puct_pre_kernel_ms'], k['puct_reroot_kernel_ms'], k['pn_kernel_ms'],
        k['puct_post_kernel_ms'], k['encode_kernel_ms'], k['network_forward_ms']))
