"""Multi-GPU parity check, run under torchrun (one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
Every rank scores its shard (halo exchange), the fused scores are all-reduced, the core-set rounds
all-gather candidate blocks; the pick list must equal the single-GPU run of the same pool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as td

import vatlq
from vatlq import dist as vd, synth


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    td.init_process_group("nccl", device_id=dev)
    comm = vd.Comm(use_p2p=os.environ.get("VATLQ_NO_P2P") is None)
    if rank == 0:
        print("candidate exchange:", "peer-memory mailbox" if comm.p2p else "ncclAllGather", flush=True)
    ok = True
    for n, k, d, n_lab, moks in ((1501, 97, 2048, 0, 0.0), (4003, 160, 2048, 300, 0.6), (777, 40, 256, 50, 0.6), (9001, 200, 2048, 900, 0.6)):
        rng = np.random.default_rng(n)
        ids, ip, inx = synth.track_flags(n, rng, mean_len=12.0)
        H = synth.heatmaps(n, seed=n, track_ids=ids)
        boxes = synth.boxes_xyxy(n, seed=n)
        X = synth.embeddings(n, d=d, seed=n + 1)
        W = synth.ae_weights(42, 4, seed=318)
        labeled = sorted(np.random.default_rng(5).choice(n, n_lab, replace=False).tolist()) if n_lab else []
        lo, hi = vd.shard_range(n, rank, world)
        res = vd.distributed_query(torch.from_numpy(H[lo:hi]).to(dev), torch.from_numpy(boxes[lo:hi]).to(dev),
                                   torch.from_numpy(ip[lo:hi].astype(np.uint8)).to(dev),
                                   torch.from_numpy(inx[lo:hi].astype(np.uint8)).to(dev),
                                   torch.from_numpy(X[lo:hi]).to(dev), W, labeled, n, k, moks, 0.01, comm=comm)
        picks = res.picks.cpu().tolist()
        one = vatlq.run_query(H, boxes, ip, inx, X, W, labeled, k, moks, 0.01, device=dev)
        same = picks == one.picks.cpu().tolist()
        thc_same = bool(torch.equal(res.thc, one.thc[lo:hi]))
        wpu_same = bool(torch.equal(res.wpu, one.wpu[lo:hi]))
        unc_same = bool(torch.equal(res.unc, one.unc))
        if rank == 0 and not same:
            a_, b_ = picks, one.picks.cpu().tolist()
            first = next((t for t in range(min(len(a_), len(b_))) if a_[t] != b_[t]), -1)
            print(f"   first divergence at pick {first}: multi {a_[first:first + 4]} vs single {b_[first:first + 4]}; "
                  f"unc_equal={unc_same} max|unc diff|={(res.unc - one.unc).abs().max().item():.3e} "
                  f"set_equal={sorted(a_) == sorted(b_)}", flush=True)
        flag = torch.tensor([int(same and thc_same and wpu_same)], device=dev)
        td.all_reduce(flag, op=td.ReduceOp.MIN)
        if rank == 0:
            print(f"n={n} k={k} d={d} labelled={n_lab}: picks_equal={same} thc_equal={thc_same} wpu_equal={wpu_same} "
                  f"all_ranks_ok={bool(flag.item())} stats={res.stats}", flush=True)
        ok = ok and bool(flag.item())
    # ---- the other strategies: sharded run vs the same code on ONE rank holding the whole pool
    solo = None
    for r in range(world):                      # (every rank must take part in every new_group call)
        g = td.new_group([r])
        if r == rank:
            solo = g
    n, k = 2203, 60
    rng = np.random.default_rng(77)
    ids, ip, inx = synth.track_flags(n, rng, mean_len=9.0)
    H = synth.heatmaps(n, seed=77, track_ids=ids)
    Hpos = np.abs(H) + np.float32(1e-3)         # (Entropy is -inf on maps with negatives: use a positive pool for it)
    boxes = synth.boxes_xyxy(n, seed=77)
    X = synth.embeddings(n, d=2048, seed=78)
    labeled = sorted(np.random.default_rng(6).choice(n, 200, replace=False).tolist())
    lo, hi = vd.shard_range(n, rank, world)
    to = lambda a, sl=slice(None): torch.from_numpy(np.ascontiguousarray(a[sl])).to(dev)
    for unc, rep, flt in (("TPC", "Influence", "None"), ("Entropy", "None", "Diversity"), ("MPE", "None", "None"),
                          ("Margin", "Influence", "Diversity"), ("HP", "Influence", "Coreset"), ("None", "Influence", "Coreset"),
                          ("THC", "None", "K-Means"), ("THC", "Influence", "weighted")):
        Hs = Hpos if unc == "Entropy" else H
        rule = "dist" if unc == "None" else "w_unc"
        kw = dict(uncertainty=unc, representativeness=rep, filter=flt, rule=rule, first_pick=-1)
        multi = vd.distributed_query(to(Hs, slice(lo, hi)), to(boxes, slice(lo, hi)), to(ip.astype(np.uint8), slice(lo, hi)),
                                     to(inx.astype(np.uint8), slice(lo, hi)), to(X, slice(lo, hi)), None, labeled, n, k, 0.6, 0.01,
                                     comm=comm, **kw)
        one = vd.distributed_query(to(Hs), to(boxes), to(ip.astype(np.uint8)), to(inx.astype(np.uint8)), to(X), None, labeled,
                                   n, k, 0.6, 0.01, comm=None, group=solo, **kw)
        same = multi.picks.cpu().tolist() == one.picks.cpu().tolist()
        close = bool(torch.allclose(multi.unc, one.unc, rtol=1e-9, atol=1e-12, equal_nan=True))
        flag = torch.tensor([int(same and close)], device=dev)
        td.all_reduce(flag, op=td.ReduceOp.MIN)
        if rank == 0:
            print(f"{unc}+{rep}_{flt}filter: picks_equal={same} scores_close={close} all_ranks_ok={bool(flag.item())}", flush=True)
        ok = ok and bool(flag.item())
    comm.close()
    td.barrier()
    td.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
