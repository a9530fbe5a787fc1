"""Short driver for ncu captures: one heat-map scan and a few core-set passes at bench scale.
    ncu --set full --clock-control none --import-source on -k regex:'scan_tma|pass_kernel' -o gpurun_out/prof python tools/prof_kernels.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vatlq
from vatlq import ops, synth

frames = int(os.environ.get("PROF_FRAMES", 20000))
rows = int(os.environ.get("PROF_ROWS", 170000))
k = int(os.environ.get("PROF_K", 40))
batch = int(os.environ.get("PROF_BATCH", 8))
dev = "cuda:0"
H, ip, inx, bb = synth.device_pool(frames, dev, seed=1)
for _ in range(2):
    r = ops.heatmap_scan(H, ip, inx, bb)
torch.cuda.synchronize()
del H
X = synth.device_embeddings(rows, dev, seed=2)
unc = torch.rand(rows, dtype=torch.float64, device=dev)
picks, st = ops.coreset_select(X, unc, [], k, 0.6, 0.01, batch=batch)
torch.cuda.synchronize()
print("ok", st)
