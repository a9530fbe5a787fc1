"""Per-round cost of the core-set loop as a function of the rows a rank owns (single GPU, clustered
counter-based pool, 5 % selected): the 8-GPU round on an N-row pool costs about what a 1-GPU round on
N/8 rows does (plus the candidate exchange), so this table is the scaling model of DESIGN.md §5.
    python tools/round_cost.py [rows ...] > gpurun_out/round_cost.json"""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vatlq
from vatlq import ops, synth

rows = [int(a) for a in sys.argv[1:]] or [21250, 42500, 85000, 125000, 170000, 250000, 500000]
dev = torch.device("cuda:0")
out = []
for n in rows:
    X = synth.pool_embeddings(n, device=dev)
    unc = synth.pool_unc(n, device=dev)
    k = min(n // 20, 8500)
    ops.coreset_select(X, unc, [], min(k, 500), 0.0, 0.01)
    torch.cuda.synchronize()
    ops.prune_stats(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    picks, st = ops.coreset_select(X, unc, [], k, 0.0, 0.01)
    e1.record()
    torch.cuda.synchronize()
    pr = ops.prune_stats(reset=True)
    ms = e0.elapsed_time(e1)
    out.append({"rows": n, "k": k, "ms": ms, "rounds": st.rounds, "passes": st.passes, "us_per_round": 1e3 * ms / st.rounds,
                "streamed": pr["streamed"] / max(1, pr["tiles"]), "mean_candidates": st.candidates / max(1, st.rounds),
                "us_tiles": st.ns_tiles / 1e3 / st.rounds, "us_plan": st.ns_plan / 1e3 / st.rounds, "us_wait": st.ns_wait / 1e3 / st.rounds})
    print(json.dumps(out[-1]), flush=True)
    del X, unc
