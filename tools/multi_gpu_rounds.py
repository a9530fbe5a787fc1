"""Per-round cost of the multi-GPU core-set loop (torchrun, one process per GPU): small shards make the
passes short, so total / rounds ~ the fixed cost of a round (exchange + pairs/plan + launches)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as td

import vatlq
from vatlq import dist as vd, ops, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
td.init_process_group("nccl", device_id=dev)
n, k = int(os.environ.get("ROWS", 16000)), int(os.environ.get("K", 2000))
X = synth.device_embeddings(n, dev, seed=2)          # same seed on every rank: replicated features
unc = torch.rand(n, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
lo, hi = vd.shard_range(n, rank, world)
for mode in ("p2p", "nccl"):
    comm = vd.Comm(use_p2p=(mode == "p2p"))
    for batch in (8, 16):
        ops.coreset_select(X, unc, [], 64, 0.6, 0.01, batch=batch, comm=comm.handle, row_range=(lo, hi))
        torch.cuda.synchronize(); td.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        picks, st = ops.coreset_select(X, unc, [], k, 0.6, 0.01, batch=batch, comm=comm.handle, row_range=(lo, hi))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0:
            print(f"{mode} batch={batch}: {dt * 1e3:.1f} ms, {st.rounds} rounds, {st.passes} passes -> {dt / st.rounds * 1e6:.1f} us/round "
                  f"(rows/rank={hi - lo})", flush=True)
    comm.close()
td.barrier()
td.destroy_process_group()
