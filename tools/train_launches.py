"""Summarise one training step out of an ncu launch list (gpu__time_duration.sum, --csv)."""
import csv, re, sys
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches_train.csv'
detail = sys.argv[2] if len(sys.argv) > 2 else None
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = [(re.sub(r'\(.*', '', x['Kernel Name']), float(x['Metric Value'].replace(',', '')), x.get('Grid Size', ''))
        for x in csv.DictReader(lines) if x.get('Metric Name') == 'gpu__time_duration.sum']
names = [r[0] for r in rows]
ia = [i for i, n in enumerate(names) if 'adam' in n]
a, b = (ia[-2] + 1, ia[-1] + 1) if len(ia) >= 2 else (0, ia[-1] + 1)
seg = rows[a:b]
tot = sum(t for _, t, _ in seg)
print('launches', len(seg), 'total ms %.3f' % (tot / 1e6))
agg = {}
for n, t, g in seg:
    k = n.split('<')[0].replace('void ', '')
    c = agg.setdefault(k, [0, 0]); c[0] += t; c[1] += 1
for k, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0])[:24]:
    print('%-50s %9.1f us %5d  %5.1f%%' % (k[:50], v / 1000, c, 100 * v / tot))
if detail:
    for n, t, g in seg:
        if detail in n:
            print('%-70s %9.1f %s' % (n[:70], t / 1000, g))
