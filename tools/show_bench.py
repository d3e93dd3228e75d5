#!/usr/bin/env python
"""Print the key figures of bench.py JSON lines (one file per argument, last line of each)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as exc:
        print(path, 'unreadable:', exc)
        continue
    r = d.get('roofline') or {}
    print('%s: %s = %.4g %s, %.4f ms/step, n_gpus %s' % (path, d.get('metric'), d.get('value') or 0, d.get('unit'),
                                                          d.get('ms_per_step') or 0, d.get('n_gpus')))
    print('   roofline %s: %.1f %s = %.3f of peak, traffic %s' % ((r.get('kernel') or '')[:50], r.get('achieved') or 0,
                                                                  r.get('unit'), r.get('frac') or 0, r.get('traffic')))
    for k in ('encoder_roofline', 'network_roofline'):
        e = d.get(k)
        if e and 'achieved' in e:
            print('   %s: %.1f %s = %.3f of peak (%.4f ms)' % (k, e['achieved'], e['unit'], e['frac'], e.get('launch_ms') or 0))
    if d.get('kernel_times_ms'):
        print('   kernels ms:', {k: round(v, 4) for k, v in d['kernel_times_ms'].items()})
    if d.get('evals_per_s'):
        print('   evals/s %.4g' % d['evals_per_s'])
    e = d.get('e2e')
    if e:
        print('   e2e %.4g %s (evals/s %s) host ms/poll %s' % (e['value'], e['unit'], e.get('evals_per_s'), e.get('host_ms_per_poll')))
    c = d.get('cpu_baseline')
    if c:
        print('   cpu_baseline %.4g %s on %s cores (%s)' % (c['value'] or 0, c['unit'], c['cores'], c['kind']))
