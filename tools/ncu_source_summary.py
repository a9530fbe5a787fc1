"""Summarise `ncu --page source --csv` output: opcode mix, stall samples per opcode, hottest lines.
    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME --launch-count 1 > src.csv
    python tools/ncu_source_summary.py src.csv
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ci, si, wi = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
tot = 0
ops, stall, reasons = collections.Counter(), collections.Counter(), collections.Counter()
data = []
for r in rows[h + 1:]:
    try:
        c, w = int(r[ci]), int(r[wi])
    except (ValueError, IndexError):
        continue
    toks = r[si].split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.split(".")[0]
    tot += c
    ops[op] += c
    stall[op] += w
    for i in stall_cols:
        try:
            reasons[hdr[i]] += int(r[i])
        except ValueError:
            pass
    data.append((c, w, r[si]))
print("warp instructions executed:", tot)
for op, c in ops.most_common(22):
    print(f"  {op:10s} {c:12d} {100 * c / tot:5.1f}%  stall samples {stall[op]}")
print("stall reasons:", ", ".join(f"{k[6:]}={v}" for k, v in reasons.most_common(8)))
print("hottest lines by stall samples:")
for c, w, s in sorted(data, key=lambda x: -x[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 18]:
    print(f"  {w:6d} {c:10d}  {s[:110]}")
