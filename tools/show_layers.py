#!/usr/bin/env python
"""Pretty-print a bench.py --layers-out file."""
import json
import sys

d = json.load(open(sys.argv[1]))
print('sum_ms', round(d['sum_ms'], 4), 'graph ms/step', round(d['graph_ms_per_step'], 4), 'batch', d['batch'])
for r in d['rows']:
    print(f"{r['name']:36s} {r['kind']:5s} {str(r['out_hw']):12s} K{r['K_tap']:4d} N{r['N']:4d} s{r['stride']} "
          f"tc={int(r['tcgen05'])} {r['ms']*1e3:8.1f}us {r['gbs']:7.0f}GB/s {r['tflops']:7.1f}TF {r['bound']:6s} "
          f"frac={r['frac']:.3f}")
