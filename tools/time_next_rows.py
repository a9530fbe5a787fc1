"""Times the SURVEY 8f kernels at bench scale (CUDA events, after warm-up): Entropy over a pool of heat
maps, cosine row sums over the feature matrix, HP/TPC from the scan outputs.  Prints achieved GB/s."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vatlq
from vatlq import ops, synth

dev = "cuda:0"
frames = int(os.environ.get("FRAMES", 40000))
rows = int(os.environ.get("ROWS", 170000))


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


out = {}
H, ip, inx, bb = synth.device_pool(frames, dev, seed=1)
t = timed(lambda: ops.heatmap_entropy(H))
out["entropy"] = {"frames": frames, "ms": t * 1e3, "GBps": frames * 17 * 64 * 48 * 4 / t / 1e9}
t = timed(lambda: ops.peak_uncertainty(H))
out["mpe_margin"] = {"frames": frames, "ms": t * 1e3, "GBps": frames * 17 * 64 * 48 * 4 / t / 1e9}
sc = torch.rand(rows, dtype=torch.float64, device=dev)
t = timed(lambda: ops.rank_scores(sc))
out["rank_scores"] = {"rows": rows, "ms": t * 1e3}
kp = torch.rand((rows, 17, 3), device=dev) * 100
t = timed(lambda: ops.oks(kp, kp + 1.0, torch.tensor([[0., 0., 99., 199.]], device=dev).repeat(rows, 1)))
out["oks"] = {"rows": rows, "ms": t * 1e3}
r = ops.heatmap_scan(H, ip, inx, bb)
t = timed(lambda: ops.pose_uncertainty(r.coords_hm, r.kpts, bb, ip, inx))
out["hp_tpc"] = {"frames": frames, "ms": t * 1e3}
del H, r
X = synth.device_embeddings(rows, dev, seed=2)
t = timed(lambda: ops.cosine_rowsum(X))
out["cosine_rowsum"] = {"rows": rows, "ms": t * 1e3, "GBps_two_passes": 2 * rows * 2048 * 4 / t / 1e9}
print(json.dumps(out))
