"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
H = rows[hdr]; data = rows[hdr + 1:]
ki = H.index('Kernel Name'); vi = H.index('Metric Value'); ui = H.index('Metric Unit')
agg = collections.OrderedDict()
for r in data:
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v)
    agg.setdefault(r[ki][:70], []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print(f"{k:72s} n={len(v):3d} total_us={sum(v):11.1f} avg_us={sum(v)/len(v):10.1f} share={sum(v)/tot*100:5.1f}%")
