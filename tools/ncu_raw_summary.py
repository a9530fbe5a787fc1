"""Compact per-launch table from an ncu report:  python tools/ncu_raw_summary.py rep.ncu-rep [> profiles/x.txt]
Columns are the ones DESIGN.md / bench.py quote (duration, DRAM bytes, DRAM %, occupancy, registers, pipes)."""
import csv
import io
import subprocess
import sys

WANT = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"),
        ("dram__bytes_write.sum", "dram_wr"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64pipe%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64cyc%"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_cyc%"),
        ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "dmma_cyc%"),
        ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tmem_inst%"),
        ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tma_cyc%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dsmem"), ("launch__shared_mem_per_block_static", "ssmem")]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
cols = [(hdr.index(k), n) for k, n in WANT if k in hdr]
if len(sys.argv) > 2 and sys.argv[2] == "--tensor":     # list every tensor / TMEM metric name the report holds
    print("# tensor-related metrics:", ", ".join(c for c in hdr if "tensor" in c or "tmem" in c))
print("# " + sys.argv[1])
print(" | ".join(f"{n}[{units[i]}]" if units[i] else n for i, n in cols))
for r in rows[2:]:
    print(" | ".join(r[i].split("(")[0][:36] if n == "kernel" else r[i] for i, n in cols))
