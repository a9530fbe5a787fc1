"""Labelled-set initialisation (min distance to the labelled rows): the exact passes (vatlq_coreset_init: n_labeled/8
pruned fp64 passes) against the tensor-core path (vatlq_coreset_init_tc: two TF32 tcgen05 GEMM sweeps + exact
re-scoring of the listed pairs).  Prints one JSON line per case.
    python tools/time_init.py [rows labelled_fraction kind] ..."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vatlq
from vatlq import ops, synth

dev = torch.device("cuda:0")
cases = [(170000, 0.1, "clustered"), (170000, 0.1, "iid"), (100000, 0.2, "weak")]
if len(sys.argv) > 3:
    cases = [(int(sys.argv[1]), float(sys.argv[2]), sys.argv[3])]
L = vatlq._lib.lib()
for n, frac, kind in cases:
    X = synth.pool_embeddings(n, kind=kind, device=dev)
    lab = torch.from_numpy(synth.pool_labeled(n, int(n * frac))).to(dev)
    md_tc = torch.empty(n, dtype=torch.float64, device=dev)
    md_ex = torch.empty(n, dtype=torch.float64, device=dev)
    ws_bytes = L.vatlq_coreset_workspace_bytes(n, 2048, 16)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)

    def exact():
        vatlq._lib.check(L.vatlq_coreset_init(C.c_void_p(X.data_ptr()), n, 2048, 0, n, C.c_void_p(lab.data_ptr()), lab.numel(),
                                              C.c_void_p(md_ex.data_ptr()), C.c_void_p(ws.data_ptr()), ws_bytes,
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)), "init")

    def tc():
        return ops.coreset_init_tc(X, lab, md_tc, 0, n)

    out = {"rows": n, "labelled": int(lab.numel()), "kind": kind}
    for name, fn in (("tc", tc), ("exact", exact)):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        out[name + "_ms"] = e0.elapsed_time(e1)
        if name == "tc":
            out["tc_ok"], out["pairs"] = bool(r[0]), r[1]["pairs"]
            out["tf32_tflops_both_sweeps"] = 2 * 2.0 * n * lab.numel() * 2048 / (out["tc_ms"] * 1e-3) / 1e12
    out["min_d_bit_equal"] = bool(torch.equal(md_tc, md_ex))
    print(json.dumps(out), flush=True)
    del X, lab, md_tc, md_ex, ws
    torch.cuda.empty_cache()
