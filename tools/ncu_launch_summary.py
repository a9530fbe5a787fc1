"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(list)
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    agg[r[ki].split("(")[0][:44]].append(v)
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':46s} launches   mean us    total ms   share")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:46s} {len(v):7d} {sum(v) / len(v):10.1f} {sum(v) / 1e3:10.2f} {100 * sum(v) / tot:6.1f}%")
