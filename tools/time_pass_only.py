"""Time ONE pass launch pattern in isolation (no selection logic): vatlq_coreset_init with 8 labelled
centres runs exactly one pass over X.  Used to compare kernel variants (VATLQ_LIB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import vatlq
from vatlq import ops, synth
lib = vatlq._lib.lib()
dev = "cuda:0"
rows = int(os.environ.get("ROWS", 170000))
X = synth.device_embeddings(rows, dev, seed=2)
lab = torch.arange(0, 8 * 1000, 1000, device=dev, dtype=torch.int64)
wsb = lib.vatlq_coreset_workspace_bytes(rows, 2048, 8)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
md = torch.empty(rows, dtype=torch.float64, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run():
    vatlq._lib.check(lib.vatlq_coreset_init(C.c_void_p(X.data_ptr()), rows, 2048, 0, rows, C.c_void_p(lab.data_ptr()), 8,
                                            C.c_void_p(md.data_ptr()), C.c_void_p(ws.data_ptr()), wsb, st))
for _ in range(3): run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print(f"{os.path.basename(os.environ.get('VATLQ_LIB','default'))}: init(8 centres) = fill + norms + pass: {e0.elapsed_time(e1)/20*1e3:.1f} us per call")
