#!/usr/bin/env python
"""bench.py — the AL query pass (THC + local-peak + WPU + fusion + core-set) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one complete query over a fixed synthetic pool (default: the PoseTrack21-sized
pool of BASELINE.json configs[3], 170 000 frames x 17 x 64x48 heat maps + 2048-d features,
5 % selected).  The pool is FIXED as N grows (strong scaling): each rank owns a contiguous
1/N of the frames.  `value` = frames of the pool / time of one query with the pool resident in
HBM; `e2e` = the same query through the public API with the pool in pinned host memory
(H2D copies of heat maps / boxes / flags / features and the D2H read of the picks inside the
timed region).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "unlabeled frames scored+selected/sec per AL query"
UNIT = "frames/s"
D = 2048
FRAME_BYTES = 17 * 64 * 48 * 4


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--frames", type=int, default=170000, help="pool size (whole job)")
    p.add_argument("--query-frac", type=float, default=0.05)
    p.add_argument("--labeled-frac", type=float, default=0.0, help="already-labelled fraction (0 = round 0)")
    p.add_argument("--moks", type=float, default=None, help="mean OKS of the last queries (default 0 at round 0, else 0.6)")
    p.add_argument("--batch", type=int, default=None,
                   help="core-set picks per round (1 = GEMV form, 8 = one pass per round, 16 = two passes per round); default 16")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-prune", action="store_true", help="core-set passes stream every tile (exact pruning off)")
    p.add_argument("--no-p2p", action="store_true", help="multi-GPU: ncclAllGather per round instead of the peer-memory mailbox")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-frames", type=int, default=512, help="frames of the CPU scoring sample")
    p.add_argument("--cpu-greedy-steps", type=int, default=6, help="greedy steps of the CPU core-set sample")
    return p.parse_args()


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled DURING the timed region: through NVML
    (nvidia_ml_py, one sample every 20 ms) when it loads, else `nvidia-smi` every 200 ms (the recipe's
    clocks line; one call takes longer than a short timed region, hence NVML first)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None
        self.source = "nvidia-smi"
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
            self.source = "nvml"
        except Exception:
            self._nvml = None

    def _run_nvml(self):
        nv = self._nvml
        masks = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
                 nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap]
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                self.rows.append([str(sm), str(self._max)] + ["Active" if r & m else "Not Active" for m in masks])
            except Exception:
                pass
            self._stop.wait(0.02)

    def _run_smi(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.rows.append(f)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run_nvml if self._nvml else self._run_smi, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"], "source": self.source}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = [nm for k, nm in enumerate(self.NAMES) if any(r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows), "source": self.source}


# ------------------------------------------------------------------------------------ CPU arm
def cpu_reference_arm(n_pool: int, k: int, moks: float, lam: float, cpu_frames: int, greedy_steps: int, seed: int = 0):
    """Time the oracle port of the reference's CPU path on a bounded sample of the workload and
    extrapolate to the whole query:  t = n/fps_scoring + k * t_greedy_step  (cost per greedy step
    is constant in the step index).  Returns (frames_per_s, detail dict)."""
    import torch
    from oracle import vatl_oracle as O
    import vatlq
    synth = vatlq.synth
    cores = os.cpu_count() or 1
    ids, ip, inx = synth.track_flags(cpu_frames, np.random.default_rng(seed), 30.0)
    H = synth.heatmaps(cpu_frames, seed=seed, track_ids=ids)
    boxes = synth.boxes_xyxy(cpu_frames, seed)
    ae = O.make_autoencoder(synth.ae_weights(42, 4))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        O.score_pool(H[:8], boxes[:8], ip[:8], inx[:8], ae)            # warm-up
        t0 = time.perf_counter()
        O.score_pool(H, boxes, ip, inx, ae)
        t_score = time.perf_counter() - t0
    fps_score = cpu_frames / t_score
    # core-set: first `greedy_steps` steps at the FULL pool size (float64 features like :270)
    X = synth.embeddings(n_pool, d=D, seed=2).astype(np.float64)
    unc = np.random.default_rng(3).uniform(0, 1, n_pool)
    O.coreset_select(X, unc.copy(), [], 1, moks, lam)                  # warm-up (BLAS threads, page faults)
    t0 = time.perf_counter()
    O.coreset_select(X, unc.copy(), [], greedy_steps, moks, lam)
    t_step = (time.perf_counter() - t0) / greedy_steps
    t_query = n_pool / fps_score + k * t_step
    detail = {"cores": cores, "torch_threads": torch.get_num_threads(), "scoring_frames_per_s": fps_score,
              "greedy_step_s": t_step, "extrapolated_query_s": t_query,
              "sample": f"scoring loop on {cpu_frames} frames + first {greedy_steps} greedy steps at N={n_pool}, "
                        f"extrapolated linearly to k={k} (t = N/fps + k*t_step)"}
    return n_pool / t_query, detail


def ncu_traffic(kernel: str, rows: int):
    """dram__bytes_read+write per launch of `kernel` from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json), valid only for the shape it was captured at; else None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[kernel]
        return t["bytes_per_launch"] if int(t["rows"]) == int(rows) else None
    except Exception:
        return None


# ------------------------------------------------------------------------------------ main
def main():
    a = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if a.batch is None:
        a.batch = 16    # two 8-centre passes per round: measured best at every GPU count once the passes are pruned
    local = int(os.environ.get("LOCAL_RANK", 0))
    n = a.frames
    n_lab = int(n * a.labeled_frac)
    k = int(n * (a.labeled_frac + a.query_frac)) - n_lab
    moks = a.moks if a.moks is not None else (0.0 if n_lab == 0 else 0.6)
    lam = 0.01
    workload = (f"full THC+WPU+core-set query, {n} frames x 17 x 64x48 fp32 heat maps + {D}-d features, "
                f"select {k} ({a.query_frac:.0%}), labelled {n_lab}, moks {moks}")
    config = {"workload": workload, "frames": n, "k": k, "feat_dim": D, "labelled": n_lab, "moks": moks,
              "unc_lambda": lam, "coreset_batch": a.batch, "parallelism": f"frame-range sharding x{world}",
              "candidate_exchange": "none" if world == 1 else ("ncclAllGather" if a.no_p2p else "peer-memory mailbox (NVLink stores + flags)"),
              "coreset_pruning": "off" if a.no_prune else "exact (segments of consecutive rows + triangle inequality; picks unchanged)",
              "l2": "inputs (>= 4 GB of heat maps + >= 174 MB of features per rank) exceed the 126 MB L2"}

    if a.impl == "reference":
        if rank != 0:
            return
        import torch  # noqa: F401
        vals = []
        det = None
        for _ in range(max(1, min(a.steps, 3))):
            v, det = cpu_reference_arm(n, k, moks, lam, a.cpu_frames, a.cpu_greedy_steps)
            vals.append(v)
        v = statistics.median(vals)
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": 1e3 * n / v, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": det["cores"], "kind": "port", "sample": det["sample"],
                                 "scoring_frames_per_s": det["scoring_frames_per_s"], "greedy_step_s": det["greedy_step_s"]},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "oracle port of the reference's CPU path (the reference is Python; /root/reference is absent on the GPU box)"}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as td
    import vatlq
    from vatlq import dist as vd, ops, synth

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    assert world == a.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    lo, hi = vd.shard_range(n, rank, world)
    nl = hi - lo

    # ---- synthetic pool shard, resident in HBM
    H, ip, inx, bb = synth.device_pool(nl, dev, seed=100 + rank)
    if world > 1:  # tracks continue across shard boundaries so that the halo frames matter
        if rank > 0:
            ip[0] = 1
        if rank < world - 1:
            inx[-1] = 1
    Xl = synth.device_embeddings(nl, dev, d=D, seed=200 + rank)
    W = synth.ae_weights(42, 4)
    gcpu = torch.Generator().manual_seed(5)
    labeled = torch.randperm(n, generator=gcpu)[:n_lab].tolist() if n_lab else []
    comm = vd.Comm(use_p2p=not a.no_p2p) if world > 1 else None
    lib = vatlq._lib.lib()
    lib.vatlq_profile_passes(1)
    if a.no_prune:
        ops.set_prune("off")

    def step_resident():
        if world == 1:
            return vatlq.run_query(H, bb, ip, inx, Xl, W, labeled, k, moks, lam, batch=a.batch, device=dev)
        return vd.distributed_query(H, bb, ip, inx, Xl, W, labeled, n, k, moks, lam, batch=a.batch, comm=comm)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
            torch.cuda.synchronize()

    for _ in range(a.warmup):
        res = step_resident()
    barrier()
    import ctypes as C
    lib.vatlq_profile_read(None, None, None, 1)
    ops.prune_stats(reset=True)
    launches0 = vatlq._lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        ev0.record()
        for _ in range(a.steps):
            res = step_resident()
        ev1.record()
        barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(ms, op=td.ReduceOp.MAX)
    ms_step = float(ms.item()) / a.steps
    launches = vatlq._lib.launch_count() - launches0
    tot_ms, n_pass, n_picks = C.c_double(), C.c_int64(), C.c_int64()
    lib.vatlq_profile_read(C.byref(tot_ms), C.byref(n_pass), C.byref(n_picks), 1)
    lib.vatlq_profile_passes(0)
    prune = ops.prune_stats(reset=True)
    streamed_frac = prune["streamed"] / prune["tiles"] if prune["tiles"] else 1.0
    st = res.stats
    picks_ref = res.picks.clone()

    # ---- dominant kernel: the pass over X (core-set)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    per_step_bytes = nl * D * 4 + 16 * nl            # one greedy step over the owned rows
    # what ONE pass must move: the rows of X it streams (all of them without pruning; the pruned tiles are
    # provably unaffected and never read) + xx r, min_d r/w, unc r/w, score w (fp64 each)  
    per_pass_bytes = nl * D * 4 * streamed_frac + 40 * nl
    roof = None
    if n_pass.value > 0 and tot_ms.value > 0:
        picks_per_launch = n_picks.value / n_pass.value
        avg_s = tot_ms.value / n_pass.value * 1e-3
        achieved = per_pass_bytes / avg_s / 1e9
        roof = {"kernel": "pass_kernel_tma (core-set distance update: TMA-staged tiles, fp64 tensor-core DMMA, row finishing)", "bound": "hbm",
                "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "traffic": ncu_traffic("pass_kernel" if a.no_prune else "pass_kernel_pruned", nl),
                "peak_source": peak_src, "avg_launch_us": avg_s * 1e6, "launches_timed": n_pass.value,
                "algorithmic_bytes_per_launch": per_pass_bytes, "streamed_fraction_of_X": streamed_frac,
                "unpruned_bytes_per_launch": nl * D * 4 + 40 * nl, "segments": prune["segments"],
                "greedy_steps_per_launch": picks_per_launch, "algorithmic_bytes_per_greedy_step": per_step_bytes,
                "greedy_equivalent_gbs": picks_per_launch * per_step_bytes / avg_s / 1e9,
                "fp64_fma_per_s": picks_per_launch * nl * D / avg_s,
                "share_of_step": tot_ms.value / a.steps / ms_step,
                "note": "frac charges a pass the bytes it moves ONCE although it applies greedy_steps_per_launch "
                        "greedy steps (exact batching); greedy_equivalent_gbs is SURVEY 8d's per-step bytes x steps / time"}
    # ---- the same kernel with pruning off (one extra, untimed-for-`value` query): how fast the pass streams
    # when it has to read every row — the pruned launches are short, so launch/prologue/epilogue weigh more
    if roof is not None and not a.no_prune and world == 1:
        unpruned = None
        try:
            ops.set_prune("off")
            lib.vatlq_profile_passes(1)
            lib.vatlq_profile_read(None, None, None, 1)
            step_resident()
            torch.cuda.synchronize()
            t2, n2, p2 = C.c_double(), C.c_int64(), C.c_int64()
            lib.vatlq_profile_read(C.byref(t2), C.byref(n2), C.byref(p2), 1)
            if n2.value > 0 and t2.value > 0:
                full = nl * D * 4 + 40 * nl
                s2 = t2.value / n2.value * 1e-3
                unpruned = {"achieved": full / s2 / 1e9, "frac": full / s2 / 1e9 / peak_gbs, "avg_launch_us": s2 * 1e6,
                            "launches_timed": n2.value, "algorithmic_bytes_per_launch": full,
                            "traffic": ncu_traffic("pass_kernel", nl)}
        except Exception as exc:  # report, never hide; the headline numbers above are unaffected
            unpruned = {"error": repr(exc)[:200]}
        finally:
            try:
                lib.vatlq_profile_passes(0)
                ops.set_prune("env")
                ops.prune_stats(reset=True)
            except Exception:
                pass
        roof["unpruned_pass"] = unpruned
        roof["note"] += ("; with exact pruning a launch streams only streamed_fraction_of_X of the rows, so its fixed "
                         "costs weigh more: unpruned_pass is the same kernel timed in one extra query with pruning off")

    # ---- the streaming kernel: heat-map scan, timed alone on this rank's shard
    scan_roof = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ops.heatmap_scan(H, ip, inx, bb)
    reps = 5
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ops.heatmap_scan(H, ip, inx, bb)
    e1.record()
    torch.cuda.synchronize()
    scan_s = e0.elapsed_time(e1) / reps * 1e-3
    scan_roof = {"kernel": "heat-map scan (THC + local peaks + argmax) + scan_finalize", "bound": "hbm",
                 "achieved": nl * FRAME_BYTES / scan_s / 1e9,
                 "peak": peak_gbs, "unit": "GB/s", "frac": nl * FRAME_BYTES / scan_s / 1e9 / peak_gbs,
                 "traffic": ncu_traffic("scan", nl),
                 "algorithmic_bytes_per_frame": FRAME_BYTES, "frames": nl, "ms": scan_s * 1e3}

    # ---- end to end: pool in pinned host memory, copies inside the timed region
    e2e = None
    if not a.no_e2e:
        try:
            e2e = run_e2e(a, vatlq, vd, H, ip, inx, bb, Xl, W, labeled, n, nl, k, moks, lam, world, comm, dev, picks_ref)
        except Exception as exc:  # report, never hide
            e2e = {"value": None, "unit": UNIT, "error": repr(exc)[:300]}

    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": n / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32 scores / f64 distance accumulation", "data": "synthetic",
                "config": config, "clocks": clk.summary(), "gpu_launches": int(launches),
                "roofline": roof, "roofline_scan": scan_roof, "e2e": e2e,
                "coreset": {"passes_over_X": st.passes, "picks": st.picks, "rounds": st.rounds,
                            "fallback_rounds": st.fallback_empty + st.fallback_overflow,
                            "mean_candidates": st.candidates / max(1, st.rounds - st.fallback_empty - st.fallback_overflow)}}
        if world == 1 and not a.no_cpu_baseline:
            try:
                v, det = cpu_reference_arm(n, k, moks, lam, a.cpu_frames, a.cpu_greedy_steps)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": det["cores"], "kind": "port",
                                        "sample": det["sample"], "scoring_frames_per_s": det["scoring_frames_per_s"],
                                        "greedy_step_s": det["greedy_step_s"]}
            except Exception as exc:
                line["cpu_baseline"] = {"value": None, "error": repr(exc)[:300]}
        print(json.dumps(line))
    if comm is not None:
        comm.close()
    if world > 1:
        td.barrier()
        td.destroy_process_group()


def run_e2e(a, vatlq, vd, H, ip, inx, bb, Xl, W, labeled, n, nl, k, moks, lam, world, comm, dev, picks_ref):
    """The same query through the public API with HOST inputs: every step copies this rank's
    heat maps, boxes, flags and features from pinned host memory and reads the picks back."""
    import torch
    import torch.distributed as td
    from vatlq import ops
    from vatlq.query import QueryPass
    chunk = 4096
    Hh = torch.empty(H.shape, dtype=H.dtype, pin_memory=True)
    for s in range(0, nl, chunk):
        Hh[s:s + chunk].copy_(H[s:s + chunk])
    Xh = torch.empty(Xl.shape, dtype=Xl.dtype, pin_memory=True); Xh.copy_(Xl)
    bbh, iph, inxh = bb.cpu().pin_memory(), ip.cpu().pin_memory(), inx.cpu().pin_memory()
    torch.cuda.synchronize()
    Hd = torch.empty_like(H)          # device landing zone (the pool as the estimator would leave it)
    Xd = torch.empty_like(Xl)
    copy_stream = torch.cuda.Stream(device=dev)
    lab = np.asarray(labeled, dtype=np.int64)
    rank = td.get_rank() if world > 1 else 0
    lo, hi = vd.shard_range(n, rank, world)

    def step():
        main = torch.cuda.current_stream()
        done = []
        with torch.cuda.stream(copy_stream):
            bd = bbh.to(dev, non_blocking=True); ipd = iph.to(dev, non_blocking=True); ind = inxh.to(dev, non_blocking=True)
            for s in range(0, nl, chunk):
                Hd[s:s + chunk].copy_(Hh[s:s + chunk], non_blocking=True)
                e = torch.cuda.Event(); e.record(copy_stream); done.append(e)
            Xd.copy_(Xh, non_blocking=True)
            ex = torch.cuda.Event(); ex.record(copy_stream)
        qp = QueryPass(nl, dev, ae_weights=W, uncertainty="THC+WPU")
        hp = hn = None
        main.wait_event(done[0])
        if world > 1:
            main.wait_event(done[-1])
            hp, hn = vd.exchange_halo(Hd[0], Hd[-1], rank, world)
        for ci, s in enumerate(range(0, nl, chunk)):
            main.wait_event(done[ci])
            e_ = min(nl, s + chunk)
            qp.score_chunk(s, Hd[s:e_], bd[s:e_], ipd[s:e_], ind[s:e_], halo_prev=hp if s == 0 else None,
                           halo_next=hn if e_ == nl else None)
        unl = torch.ones(nl, dtype=torch.uint8, device=dev)
        mine = lab[(lab >= lo) & (lab < hi)] - lo
        if mine.size:
            unl[torch.from_numpy(mine).to(dev)] = 0
        unc_l = qp.fuse(unl, "const", labeled_ratio=lab.size / n, group=None if world == 1 else td.group.WORLD,
                        n_unlabeled_global=n - lab.size)
        main.wait_event(ex)
        X = vd.allgather_rows(Xd, n, world) if world > 1 else Xd
        unc = vd.allgather_rows(unc_l, n, world) if world > 1 else unc_l
        picks, _ = ops.coreset_select(X, unc, lab, k, moks, lam, batch=a.batch,
                                      comm=comm.handle if comm is not None else None,
                                      row_range=(lo, hi) if world > 1 else None)
        return picks.cpu()      # D2H read of the result

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
            torch.cuda.synchronize()

    p = step()
    same = bool(torch.equal(p, picks_ref.cpu()))
    reps = max(1, min(a.steps, 3))
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        p = step()
    sync()
    dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(dt, op=td.ReduceOp.MAX)
    h2d = nl * FRAME_BYTES + nl * (16 + 2) + nl * D * 4
    return {"value": n / float(dt.item()), "unit": UNIT, "ms_per_step": float(dt.item()) * 1e3,
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(k * 8), "steps": reps,
            "picks_equal_resident_run": same,
            "api": "vatlq.QueryPass.score_chunk + fuse + ops.coreset_select (host pinned inputs, per rank)"}


if __name__ == "__main__":
    main()
