#!/usr/bin/env python
"""Prints a compact summary of a bench.py JSON line (file argument or stdin)."""
import json
import sys

txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
lines = [l for l in txt.splitlines() if l.startswith('{')]
if not lines:
    print(txt[-3000:])
    sys.exit(1)
d = json.loads(lines[-1])
print('C2 value %.0f Mpix/s  ms/step %.4f  e2e %.0f Mpix/s  launches/step %s  n_gpus %s' % (
    d['value'], d['ms_per_step'], d['e2e']['value'], d.get('launches_per_step'), d['n_gpus']))
r = d.get('roofline')
if r:
    print('  %s: %.1f us  %.0f GB/s  frac %.3f' % (r.get('kernel'), r['ms_per_launch'] * 1e3, r['achieved'], r['frac']))
if 'cpu_baseline' in d:
    print('  cpu %.1f Mpix/s on %s cores; clocks %s' % (d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d.get('clocks')))
    a = d['cpu_baseline'].get('aten_chain_same_gpu')
    if a and 'value' in a:
        print('  ATen chain on the same GPU: %.0f Mpix/s (%.2f ms/step)' % (a['value'], a['ms_per_step']))
for k, v in d.get('workloads', {}).items():
    if isinstance(v, dict) and 'fwd' in v:
        fr = lambda part: ('%.2f' % part['roofline']['frac']) if 'roofline' in part else 'n/a'
        print('  %-28s fwd %.3f ms (%s)  fwd+bwd %.3f ms (%s)  %s' % (
            k, v['fwd']['ms'], fr(v['fwd']), v['fwd_bwd']['ms'], fr(v['fwd_bwd']),
            (v['fwd_bwd'].get('plan') or v.get('plan', ''))[:46]))
    elif isinstance(v, dict) and 'ms' in v:
        print('  %-28s %.3f ms  %.0f Mpix/s  frac %.2f' % (k, v['ms'], v['mpix_s'], v['roofline']['frac']))
    else:
        print('  ', k, v)
