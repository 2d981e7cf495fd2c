"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: python scripts/launch_summary.py launches.csv"""
import csv, collections, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg, tot = collections.OrderedDict(), 0.0
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    k = re.sub(r'\(.*', '', row['Kernel Name'])[:60]
    v = float(row['Metric Value'].replace(',', ''))
    v = v / 1000 if row['Metric Unit'] == 'ns' else v * 1000 if row['Metric Unit'] == 'ms' else v
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(f'total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches')
for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f'{k:60s} n={n:4d} total={v:9.1f}us avg={v/n:8.2f}us share={100*v/tot:5.1f}%')
