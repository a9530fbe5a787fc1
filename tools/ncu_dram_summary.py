"""Per-kernel totals from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`
launch list: launches, mean duration, mean DRAM bytes per launch, share of the summed GPU time."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
dur, rd, wr = collections.defaultdict(list), collections.defaultdict(float), collections.defaultdict(float)
for r in rows[1:]:
    name = r[ki].split("(")[0][:40]
    v = float(r[vi].replace(",", "")) * SCALE.get(r[ui], 1.0)
    if r[mi].startswith("gpu__time_duration"):
        dur[name].append(v)
    elif r[mi].startswith("dram__bytes_read"):
        rd[name] += v
    elif r[mi].startswith("dram__bytes_write"):
        wr[name] += v
tot = sum(sum(v) for v in dur.values())
print(f"{'kernel':42s} launches   mean us    total ms   share   dram rd MB/launch   dram wr MB/launch")
for k, v in sorted(dur.items(), key=lambda kv: -sum(kv[1])):
    n = len(v)
    print(f"{k:42s} {n:7d} {sum(v) / n:10.1f} {sum(v) / 1e3:10.2f} {100 * sum(v) / tot:6.1f}% {rd[k] / n / 1e6:16.2f} {wr[k] / n / 1e6:18.2f}")
