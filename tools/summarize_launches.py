#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for row in r:
    if len(row) <= vi: continue
    v = float(row[vi].replace(",", ""))
    u = row[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    name = re.sub(r"\(.*", "", row[ki])[:70]
    agg[name][0] += 1; agg[name][1] += ms
tot = sum(v[1] for v in agg.values())
print("total %.3f ms over %d launches" % (tot, sum(v[0] for v in agg.values())))
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%6.2f%% %6d launches %10.3f ms  %s" % (100 * ms / tot, n, ms, k))
