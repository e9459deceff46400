#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and mean time per kernel."""
import csv, sys, collections, re
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik]); name = re.sub(r"^void |b2::", "", name)
    v = float(r[iv].replace(",", "")); u = r[iu]
    v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(u, 1.0)
    tot[name] += v; cnt[name] += 1
allt = sum(tot.values())
print(f"{'kernel':48s} {'launches':>8s} {'total us':>10s} {'mean us':>8s} {'share':>6s}")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"{k[:48]:48s} {cnt[k]:8d} {tot[k]:10.1f} {tot[k]/cnt[k]:8.2f} {100*tot[k]/allt:5.1f}%")
print(f"{'total':48s} {sum(cnt.values()):8d} {allt:10.1f}")
