"""Summarise an `ncu --csv --metrics ...` log: one record per (kernel, grid, block) with the mean of every metric.
Usage: python tools/ncu_csv_summary.py gpurun_out/x.csv [out.json]"""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    name = r[ix["Kernel Name"]]
    short = name.split("(")[0].split("::")[-1]
    k = (short, r[ix["Grid Size"]], r[ix["Block Size"]])
    d = agg.setdefault(k, collections.OrderedDict())
    try:
        val = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    d.setdefault(r[ix["Metric Name"]], []).append(val)
out = []
for (name, grid, block), d in agg.items():
    rec = collections.OrderedDict(kernel=name, grid=grid, block=block, launches=max(len(v) for v in d.values()))
    for m, v in d.items():
        rec[m] = round(sum(v) / len(v), 3)
    out.append(rec)
js = json.dumps(out, indent=1)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(js + "\n")
else:
    print(js)
