"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
cur = None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if r[0] in ('Function Name', 'Line No'):
        continue
    if len(r) > 8 and r[2] == '-':
        try:
            line = int(r[0]); samples = int(r[4]); inst = int(r[7])
        except ValueError:
            continue
        k = (cur, line)
        a = agg.setdefault(k, [0, 0, r[1]])
        a[0] += samples; a[1] += inst
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print('total samples', ts, 'warp instructions', ti)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:4d} samp {100*v[0]/ts:5.1f}% inst {100*v[1]/ti:5.1f}%  {v[2][:120]}")
