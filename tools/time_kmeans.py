"""Times the device K-Means filter (kmeans.py / csrc/kmeans.cu) and, on a bounded size, sklearn's KMeans on the host
cores next to it.  CUDA events around the whole call and around the two GEMM-shaped kernels; prints one JSON line.
    python tools/time_kmeans.py            # SIZES="n:k,..." CPU="n:k" to override"""
import ctypes as C
import json
import os
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import vatlq
from vatlq import kmeans as KM, synth

dev = "cuda:0"
sizes = [tuple(int(v) for v in s.split(":")) for s in os.environ.get("SIZES", "3000:150,20000:1000,100000:5000").split(",")]
cpu_n, cpu_k = (int(v) for v in os.environ.get("CPU", "10000:500").split(":"))
L = vatlq._lib.lib()
p = lambda t: C.c_void_p(t.data_ptr())
out = {"device": torch.cuda.get_device_name(0), "runs": []}


def ev():
    return torch.cuda.Event(enable_timing=True)


for n, k in sizes:
    X = synth.pool_embeddings(n, kind="weak", seed=2, device=dev)
    d = X.shape[1]
    KM.kmeans_fit_select(X[:2000], 20)                       # warm-up (module load, attribute set-up)
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    res = KM.kmeans_fit_select(X, k)
    e1.record(); torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1)
    # the assignment GEMM alone (rows x centres x d on the fp64 tensor cores)
    ws = torch.empty(int(L.vatlq_kmeans_workspace_bytes(n, d, k)), dtype=torch.uint8, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    labels = torch.empty(n, dtype=torch.int32, device=dev)
    L.vatlq_kmeans_assign(p(X), n, d, p(res.centers), k, p(labels), None, None, p(ws), ws.numel(), st)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        L.vatlq_kmeans_assign(p(X), n, d, p(res.centers), k, p(labels), None, None, p(ws), ws.numel(), st)
    e1.record(); torch.cuda.synchronize()
    assign_ms = e0.elapsed_time(e1) / 3
    # the seeding alone
    cid = torch.empty(k, dtype=torch.int32, device=dev)
    closest = torch.empty(n, dtype=torch.float64, device=dev)
    trials = 2 + int(np.log(k))
    rand = torch.rand((k - 1) * trials, dtype=torch.float64, device=dev)
    e0.record()
    L.vatlq_kmeans_pp(p(X), n, d, None, k, 0, p(rand), trials, p(cid), p(closest), p(ws), ws.numel(), st)
    e1.record(); torch.cuda.synchronize()
    pp_ms = e0.elapsed_time(e1)
    fma = n * k * d
    out["runs"].append({"n": n, "k": k, "d": d, "total_ms": total_ms, "n_iter": res.n_iter, "relocations": res.relocations,
                        "assign_ms": assign_ms, "assign_fp64_tflops": 2 * fma / (assign_ms * 1e-3) / 1e12,
                        "pp_ms": pp_ms, "pp_us_per_step": pp_ms * 1e3 / k,
                        "pp_x_read_GBps": k * n * d * 4 / (pp_ms * 1e-3) / 1e9, "trials": trials})
    del X, ws
    torch.cuda.empty_cache()

# sklearn on the host cores, same pool construction, bounded size
from sklearn.cluster import KMeans
Xh = synth.pool_embeddings(cpu_n, kind="weak", seed=2).astype(np.float64)
t0 = time.time()
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    km = KMeans(n_clusters=cpu_k, random_state=318).fit(Xh)
cpu_s = time.time() - t0
Xd = torch.from_numpy(Xh.astype(np.float32)).to(dev)
torch.cuda.synchronize()
t0 = time.time()
res = KM.kmeans_fit_select(Xd, cpu_k)
torch.cuda.synchronize()
gpu_s = time.time() - t0
out["cpu"] = {"n": cpu_n, "k": cpu_k, "sklearn_s": cpu_s, "sklearn_iter": int(km.n_iter_), "device_s": gpu_s, "device_iter": res.n_iter,
              "labels_equal": bool(np.array_equal(res.labels.cpu().numpy(), km.labels_)), "cores": os.cpu_count()}
print(json.dumps(out))
