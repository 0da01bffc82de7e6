"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, collections, re, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', row['Kernel Name'])[:80]
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    ms = v / 1e6 if unit in ('ns', 'nsecond') else (v / 1e3 if unit in ('us', 'usecond') else v)
    agg[name][0] += 1
    agg[name][1] += ms
    tot += ms
print('total ms', round(tot, 3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{v[1]:9.3f} x{v[0]:<5d} {k}")
