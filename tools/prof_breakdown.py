import json, sys
d=json.load(open(sys.argv[1]))
pat = sys.argv[2] if len(sys.argv)>2 else ''
rows=sorted(d['detail'].items(), key=lambda kv:-kv[1]['ms_per_step'])
tot=sum(v['ms_per_step'] for k,v in rows)
for k,v in rows:
    if pat in k: print(f"{v['ms_per_step']:8.3f} x{v['calls_per_step']:<4.0f} {k}")
print('total', tot)
