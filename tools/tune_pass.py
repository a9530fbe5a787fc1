"""Time the core-set pass kernel for one library build (VATLQ_LIB) at two shard sizes."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vatlq
from vatlq import ops, synth

lib = vatlq._lib.lib()
dev = "cuda:0"
BATCH = int(os.environ.get("BATCH", 8))
for rows, k in ((170000, 800), (21250, 800)):  # run under `timeout -s KILL`
    X = synth.device_embeddings(rows, dev, seed=2)
    unc = torch.rand(rows, dtype=torch.float64, device=dev)
    ops.coreset_select(X, unc, [], 16, 0.6, 0.01, batch=BATCH)
    lib.vatlq_profile_passes(1)
    lib.vatlq_profile_read(None, None, None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    picks, st = ops.coreset_select(X, unc, [], k, 0.6, 0.01, batch=BATCH)
    e1.record()
    torch.cuda.synchronize()
    ms, n, p = C.c_double(), C.c_int64(), C.c_int64()
    lib.vatlq_profile_read(C.byref(ms), C.byref(n), C.byref(p), 1)
    lib.vatlq_profile_passes(0)
    us = ms.value / max(n.value, 1) * 1e3
    print(f"{os.path.basename(os.environ.get('VATLQ_LIB', 'default'))}: rows={rows} pass={us:.1f} us  read={rows * 2048 * 4 / us / 1e3:.0f} GB/s "
          f"fma={rows * 2048 * 8 / us / 1e6:.2f} T/s  total={e0.elapsed_time(e1):.1f} ms for {k} picks ({st.passes} passes)\n   {st}")
    del X
