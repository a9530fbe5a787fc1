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


def _kmeans_worker(rank, world, port, q):
    """kmeans.py with group=WORLD on two CPU processes: the assignment step is sharded by rows, labels all-gathered, the
    change count summed; the device calls are the NumPy stand-ins of tests/test_kmeans_host_cpu.py."""
    import contextlib
    import sys
    import warnings
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [os.path.dirname(here), here]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import vatlq  # noqa: F401
        from vatlq import _lib, kmeans as KM
        from test_kmeans_host_cpu import FakeLib
        from sklearn.cluster import KMeans
        fake = FakeLib()
        _lib.lib = lambda: fake
        KM._cuda = lambda t, dt, name: t.contiguous()
        KM._stream = lambda: None
        torch.cuda.device = lambda dev: contextlib.nullcontext()
        rng = np.random.default_rng(4)                       # identical pool on every rank
        n, d, k = 301, 20, 23                                # ragged shards
        X = np.abs(rng.normal(0, 1, (n, d))).astype(np.float32)
        w = 1 + rng.random(n)
        fake.n, fake.k = n, k
        res = KM.kmeans_fit_select(torch.from_numpy(X), k, sample_weight=torch.from_numpy(w), group=td.group.WORLD)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            km = KMeans(n_clusters=k, random_state=318)
            lab = km.fit_predict(X.astype(np.float64), sample_weight=w)
        ok = bool(np.array_equal(res.labels.numpy(), lab)) and res.n_iter == km.n_iter_
        rows = torch.tensor(res.query_rows)
        both = [torch.empty_like(rows) for _ in range(world)]
        td.all_gather(both, rows)
        ok &= all(torch.equal(b, rows) for b in both)        # every rank returns the same picks
        q.put((rank, ok))
    finally:
        td.destroy_process_group()


def test_kmeans_row_sharded_assignment_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_kmeans_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)], res
