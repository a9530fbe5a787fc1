"""CPU, world_size 2, gloo: the host-side plumbing of the multi-GPU path (halo exchange,
ragged all-gather, MIN all-reduce of the fusion statistics)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as td
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vatlq import dist
        g = torch.Generator().manual_seed(7)
        H = torch.randn((n, 3, 4, 4), generator=g)          # identical on every rank
        lo, hi = dist.shard_range(n, rank, world)
        loc = H[lo:hi]
        hp, hn = dist.exchange_halo(loc[0], loc[-1], rank, world)
        ok = True
        ok &= (hp is None) if lo == 0 else torch.equal(hp, H[lo - 1])
        ok &= (hn is None) if hi == n else torch.equal(hn, H[hi])
        full = dist.allgather_rows(loc.reshape(hi - lo, -1), n, world)
        ok &= torch.equal(full, H.reshape(n, -1))
        # 1-D int32 shards (the K-Means labels of the row-sharded assignment step, kmeans.py) + SUM of the change counts
        lab = (torch.arange(n, dtype=torch.int32) * 7) % 5
        ok &= torch.equal(dist.allgather_rows(lab[lo:hi], n, world), lab)
        cnt = torch.tensor([hi - lo], dtype=torch.int32)
        td.all_reduce(cnt, op=td.ReduceOp.SUM)
        ok &= int(cnt.item()) == n
        # fusion statistics: {min, -max} pairs reduce with one MIN
        v = H[lo:hi].reshape(-1).double()
        st = torch.stack([v.min(), (-v).min()])
        td.all_reduce(st, op=td.ReduceOp.MIN)
        ok &= bool(st[0] == H.min().double()) and bool(-st[1] == H.max().double())
        q.put((rank, bool(ok)))
    finally:
        td.destroy_process_group()


def test_halo_and_gather_world2():
    for n in (9, 10):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        ps = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
        for p in ps:
            p.start()
        res = [q.get(timeout=120) for _ in ps]
        for p in ps:
            p.join(timeout=60)
        assert sorted(res) == [(0, True), (1, True)], res
