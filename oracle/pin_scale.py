#!/usr/bin/env python
"""Pin the core-set selection at BASELINE scale: run the REFERENCE's own
`ActiveLearning.coreset_selection` (active_learning/ActiveLearning.py:798-850, imported from
/root/reference with the stub recipe of pin_against_reference.py) on the exact pools bench.py and the
`-m gpu` scale tests use (the counter-based generators of vatl4pose-wacv2024_b200/synth.py produce the
same bits in numpy here and in torch on the GPU), and store ONLY the pick lists + their sha256 under
tests/golden/coreset_scale_<tag>.npz.  TEST INFRASTRUCTURE; runs only in the build container.

    python oracle/pin_scale.py c3           # config 3: 100 000 x 2048, k = 5 000, round 0          (~35 min, 8 cores)
    python oracle/pin_scale.py c3lab        # config 3 pool, 10 % labelled, moks 0.6, first 400 picks
    python oracle/pin_scale.py c4           # config 4: full THC+WPU+core-set query, 170 000 frames, k = 8 500 (~2 h)
    python oracle/pin_scale.py c4lab        # config 4 pool, 10 % labelled, moks 0.6, first 400 picks
    python oracle/pin_scale.py c5           # config 5: 1 M frames, full scoring + the first 400 picks of k = 50 000
    python oracle/pin_scale.py weak|iid     # control pools (3:1 separation / i.i.d. rows), 20 000 rows, k = 1 000

"full query" = heat maps -> oracle scoring loop (vatl_oracle.score_pool, itself pinned bit for bit against the
reference) -> fusion -> the reference's coreset_selection.  Scoring is spread over worker processes.
"""
from __future__ import annotations

import hashlib
import importlib
import io
import contextlib
import os
import sys
import time
import warnings
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
LAM = 0.01


def sha(picks) -> str:
    return hashlib.sha256(np.asarray(picks, dtype="<i8").tobytes()).hexdigest()


def _score_chunk(args):
    """worker: oracle scores of pool items [a, b) (neighbour frames regenerated locally)"""
    n, a, b, seed = args
    from oracle import vatl_oracle as O
    synth = importlib.import_module("vatl4pose-wacv2024_b200.synth")
    tid, pos, ip, inx = _score_chunk.tracks
    lo, hi = max(0, a - 1), min(n, b + 1)
    H = synth.pool_heatmaps(tid, pos, lo, hi, seed=seed)
    boxes = synth.pool_boxes(n, lo, hi, seed=seed)
    ae = O.make_autoencoder(synth.ae_weights(42, 4))
    fp, fn = ip[lo:hi].copy(), inx[lo:hi].copy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = O.score_pool(H, boxes, fp, fn, ae)
    s = slice(a - lo, b - lo)
    return a, b, r["thc"][s], r["wpu"][s], r["peak"][s]


def _init_worker(tracks):
    _score_chunk.tracks = tracks
    try:
        import torch
        torch.set_num_threads(1)
    except Exception:
        pass


def full_scores(n: int, seed: int, workers: int):
    """THC / WPU / peak mean of every pool item through the oracle's per-person loop."""
    import multiprocessing as mp
    synth = importlib.import_module("vatl4pose-wacv2024_b200.synth")
    tracks = synth.pool_tracks(n, seed=seed)
    thc, wpu, peak = np.zeros(n), np.zeros(n), np.zeros(n)
    step = 500
    jobs = [(n, a, min(n, a + step), seed) for a in range(0, n, step)]
    t0 = time.time()
    with mp.get_context("fork").Pool(workers, initializer=_init_worker, initargs=(tracks,)) as pool:
        for q, (a, b, t, w, p) in enumerate(pool.imap_unordered(_score_chunk, jobs)):
            thc[a:b], wpu[a:b], peak[a:b] = t, w, p
            if q % 50 == 0:
                print(f"  scored {q * step}/{n}  {time.time() - t0:.0f} s", flush=True)
    return thc, wpu, peak


def reference_select(X64, unc, labeled, k, moks):
    from pin_against_reference import import_reference
    R = import_reference()
    fake = SimpleNamespace(labeled_id=R.IndexCollection([int(i) for i in labeled]), moks_queried=moks, unc_lambda=LAM,
                           uncertainty="THC+WPU", cfg=SimpleNamespace(VAL=SimpleNamespace(UNC_LAMBDA=LAM)),
                           opt=SimpleNamespace(fixed_lambda=False), query_size=k)
    import sklearn
    # assume_finite only skips sklearn's NaN/inf validation pass over X in every greedy step (no arithmetic changes)
    with contextlib.redirect_stderr(io.StringIO()), sklearn.config_context(assume_finite=True):
        return [int(i) for i in R.AL.coreset_selection(fake, X64, unc)]


def main():
    tag = sys.argv[1]
    workers = int(os.environ.get("PIN_WORKERS", "4"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    synth = importlib.import_module("vatl4pose-wacv2024_b200.synth")
    from oracle import vatl_oracle as O
    cfg = {
        "c3": dict(n=100000, k=5000, lab=0.0, moks=0.0, kind="clustered", full=False, k_full=5000),
        "c3lab": dict(n=100000, k=400, lab=0.1, moks=0.6, kind="clustered", full=False, k_full=5000),
        "c4": dict(n=170000, k=8500, lab=0.0, moks=0.0, kind="clustered", full=True, k_full=8500),
        "c4lab": dict(n=170000, k=400, lab=0.1, moks=0.6, kind="clustered", full=True, k_full=8500),
        "c5": dict(n=1000000, k=400, lab=0.0, moks=0.0, kind="clustered", full=True, k_full=50000),
        "weak": dict(n=20000, k=1000, lab=0.0, moks=0.6, kind="weak", full=False, k_full=1000),
        "iid": dict(n=20000, k=1000, lab=0.0, moks=0.6, kind="iid", full=False, k_full=1000),
    }[tag]
    n, k = cfg["n"], cfg["k"]
    t0 = time.time()
    lab = synth.pool_labeled(n, int(n * cfg["lab"]))
    extra = {}
    if cfg["full"]:
        cache = f"/tmp/pin_scale_scores_{n}.npz"
        if os.path.exists(cache):
            z = np.load(cache)
            thc, wpu, peak = z["thc"], z["wpu"], z["peak"]
        else:
            thc, wpu, peak = full_scores(n, 0, workers)
            np.savez(cache, thc=thc, wpu=wpu, peak=peak)
        unl = np.ones(n, dtype=bool)
        unl[lab] = False
        unc = np.zeros(n)
        unc[unl] = O.fuse_scores(thc[unl], wpu[unl], "const", lab.size / n)          # ActiveLearning.py:490-516,609-611
        top = np.argsort(-unc, kind="stable")[:4]
        extra = dict(top_unc_idx=top.astype(np.int64), top_unc=unc[top], thc_sum=np.float64(thc.sum()),
                     wpu_sum=np.float64(wpu.sum()), combine_weight=np.float64(np.nansum(peak[unl]) / unl.sum()),
                     thc_head=thc[:64].copy(), wpu_head=wpu[:64].copy(), peak_head=peak[:64].copy())
        print(f"scores done {time.time() - t0:.0f} s; top unc {top[:2]} {unc[top[:2]]}", flush=True)
    else:
        unc = synth.pool_unc(n)
        unc[lab] = 0.0
    X = synth.pool_embeddings(n, kind=cfg["kind"]).astype(np.float64)                 # ActiveLearning.py:270,286
    print(f"pool ready {time.time() - t0:.0f} s", flush=True)
    t1 = time.time()
    picks = reference_select(X, unc.copy(), lab, k, cfg["moks"])
    dt = time.time() - t1
    print(f"{tag}: {k} picks in {dt:.0f} s ({dt / k:.3f} s/step), sha256 {sha(picks)}", flush=True)
    np.savez_compressed(os.path.join(GOLD, f"coreset_scale_{tag}.npz"), picks=np.asarray(picks, np.int64), n=n, k=k,
                        k_full=cfg["k_full"], labeled_frac=cfg["lab"], moks=cfg["moks"], lam=LAM, kind=cfg["kind"],
                        full_query=cfg["full"], sha256=sha(picks), seconds=dt, **extra)


if __name__ == "__main__":
    main()
