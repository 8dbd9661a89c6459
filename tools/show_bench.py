#!/usr/bin/env python
"""Pretty-print a bench.py JSON line (per-kernel table + roofline)."""
import json, signal, sys
signal.signal(signal.SIGPIPE, signal.SIG_DFL)
for l in open(sys.argv[1]):
    if not l.startswith('{'): continue
    d = json.loads(l)
    print(f"ms/step {d['ms_per_step']:.4f}  value {d['value']:.4g} {d['unit']}  e2e {d['e2e']['value']:.4g} ({d['e2e']['ms_per_step']:.2f} ms)  launches {d['gpu_launches']}  clocks {d['clocks']}")
    for k, v in sorted(d['kernels'].items(), key=lambda kv: -kv[1]['ms_per_step']):
        print(f"  {k:28s} {v['ms_per_step']*1e3:8.1f} us  x{v['launches_per_step']:.0f}  {v['share']:6.1%}")
    r = d['roofline']
    print('dominant', r['kernel'], f"frac {r['frac']:.4f}", {k: round(v['frac'], 4) for k, v in r['contract_kernels'].items()}, 'whole step frac', round(r['whole_step']['frac'], 4))
    print('cpu', d['cpu_baseline'])
