#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    a = agg.setdefault(r[kn].split("(")[0][-50:], [0, 0.0]); a[0] += 1; a[1] += float(r[mv].replace(",", ""))
tot = sum(a[1] for a in agg.values())
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{n:50s} n={c:4d} avg={t/c/1e3:8.1f} us share={t/tot*100:5.1f}%")
