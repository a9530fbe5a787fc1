"""Seeded synthetic inputs of PoseTrack21 / SimpleBaseline shape.

Everything here is host-side NumPy (deterministic for a given seed); `device_pool`
builds the same kind of data directly in HBM with torch for pools that are too
large to stage through the host.  Shapes follow the reference config
`configs/posetrack21/al_simple_posetrack.yaml:21-28` (17 joints, 64x48, sigma 2).
"""
from __future__ import annotations

import numpy as np

J, HM_H, HM_W, FEAT_D = 17, 64, 48, 2048
FRAME_BYTES = J * HM_H * HM_W * 4


def track_flags(n: int, rng: np.random.Generator, mean_len: float = 30.0):
    """Track ids with geometric lengths; isPrev/isNext exactly as the dataset derives
    them from track_id equality (reference alphapose/datasets/posetrack21.py:148-178)."""
    ids = np.empty(n, dtype=np.int64)
    i, t = 0, 0
    while i < n:
        ln = int(rng.geometric(1.0 / mean_len))
        ids[i:i + ln] = t
        i += ln
        t += 1
    is_prev = np.zeros(n, dtype=np.uint8)
    is_next = np.zeros(n, dtype=np.uint8)
    if n > 1:
        same = ids[1:] == ids[:-1]
        is_prev[1:] = same
        is_next[:-1] = same
    return ids, is_prev, is_next


def heatmaps(n: int, seed: int = 0, track_ids=None) -> np.ndarray:
    """(n,17,64,48) fp32: per joint a sigma=2 blob on a random walk (step <= 1.5 px per
    frame, restarted at track starts), amplitude U(0.3,1), N(0,0.02) noise, and with
    p=0.1 a secondary blob of 0.4-0.9x amplitude."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:HM_H, 0:HM_W].astype(np.float32)
    out = np.empty((n, J, HM_H, HM_W), dtype=np.float32)
    cx = rng.uniform(6, HM_W - 6, size=J)
    cy = rng.uniform(6, HM_H - 6, size=J)
    for i in range(n):
        if track_ids is not None and i > 0 and track_ids[i] != track_ids[i - 1]:
            cx = rng.uniform(6, HM_W - 6, size=J)
            cy = rng.uniform(6, HM_H - 6, size=J)
        else:
            cx = np.clip(cx + rng.uniform(-1.5, 1.5, size=J), 2, HM_W - 3)
            cy = np.clip(cy + rng.uniform(-1.5, 1.5, size=J), 2, HM_H - 3)
        amp = rng.uniform(0.3, 1.0, size=J).astype(np.float32)
        g = amp[:, None, None] * np.exp(
            -((xx[None] - cx[:, None, None].astype(np.float32)) ** 2
              + (yy[None] - cy[:, None, None].astype(np.float32)) ** 2) / 8.0).astype(np.float32)
        sec = rng.random(J) < 0.1
        if sec.any():
            sx = rng.uniform(3, HM_W - 4, size=J).astype(np.float32)
            sy = rng.uniform(3, HM_H - 4, size=J).astype(np.float32)
            sa = (amp * rng.uniform(0.4, 0.9, size=J).astype(np.float32)) * sec
            g = g + sa[:, None, None] * np.exp(
                -((xx[None] - sx[:, None, None]) ** 2 + (yy[None] - sy[:, None, None]) ** 2) / 8.0
            ).astype(np.float32)
        out[i] = g + rng.normal(0.0, 0.02, size=g.shape).astype(np.float32)
    return out


def boxes_xyxy(n: int, seed: int = 0) -> np.ndarray:
    """(n,4) fp32 crop boxes: centre U(100,1000)xU(100,700), height U(80,400), aspect 0.75."""
    rng = np.random.default_rng(seed + 7919)
    cx = rng.uniform(100, 1000, n)
    cy = rng.uniform(100, 700, n)
    h = rng.uniform(80, 400, n)
    w = 0.75 * h
    return np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], axis=1).astype(np.float32)


def poses(n: int, seed: int = 1):
    """Config-2 style poses: box-relative skeleton + N(0,5 px), scores U(0.1,1).
    Returns keypoints (n,17,3) fp32 [x,y,score] and boxes (n,4) fp32 xyxy."""
    rng = np.random.default_rng(seed)
    bb = boxes_xyxy(n, seed)
    # a crude upright skeleton in box-normalised coordinates (COCO joint order)
    sk = np.array([[.5, .08], [.54, .06], [.46, .06], [.58, .08], [.42, .08], [.65, .22], [.35, .22],
                   [.72, .38], [.28, .38], [.74, .52], [.26, .52], [.60, .55], [.40, .55],
                   [.62, .75], [.38, .75], [.63, .95], [.37, .95]])
    w = (bb[:, 2] - bb[:, 0])[:, None]
    h = (bb[:, 3] - bb[:, 1])[:, None]
    x = bb[:, 0:1] + sk[None, :, 0] * w + rng.normal(0, 5.0, (n, J))
    y = bb[:, 1:2] + sk[None, :, 1] * h + rng.normal(0, 5.0, (n, J))
    s = rng.uniform(0.1, 1.0, (n, J))
    return np.stack([x, y, s], axis=2).astype(np.float32), bb


def embeddings(n: int, d: int = FEAT_D, seed: int = 2, clustered: bool = True) -> np.ndarray:
    """(n,d) fp32 pooled-feature stand-ins.  clustered: ceil(n/30) centres ~ ReLU(N(0,1))*0.5
    plus N(0,0.01) within-cluster noise (temporal clusters, the fp32-vs-fp64 hazard of
    SURVEY.md §7.3-1); otherwise i.i.d. ReLU(N(0,1))."""
    rng = np.random.default_rng(seed)
    if not clustered:
        return np.maximum(rng.standard_normal((n, d)), 0).astype(np.float32)
    nc = -(-n // 30)
    cen = (np.maximum(rng.standard_normal((nc, d)), 0) * 0.5).astype(np.float32)
    assign = np.minimum(np.arange(n) // 30, nc - 1)
    return (cen[assign] + rng.normal(0, 0.01, (n, d)).astype(np.float32)).astype(np.float32)


def ae_weights(in_dim: int = 42, z_dim: int = 4, seed: int = 318):
    """Random-init WholeBodyAE parameters with torch's default nn.Linear init, as a list of
    (W[out,in], b[out]) fp32 arrays for the 8 layers in->24->12->7->z->7->12->24->in
    (reference active_learning/Whole_body_AE/AutoEncoder.py:13-32)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    dims = [in_dim, 24, 12, 7, z_dim, 7, 12, 24, in_dim]
    out = []
    for a, b in zip(dims[:-1], dims[1:]):
        bound = 1.0 / np.sqrt(a)
        W = (torch.rand((b, a), generator=g) * 2 - 1) * bound
        bias = (torch.rand((b,), generator=g) * 2 - 1) * bound
        out.append((W.numpy().astype(np.float32), bias.numpy().astype(np.float32)))
    return out


def device_pool(n: int, device, seed: int = 0, chunk: int = 4096, mean_len: float = 30.0):
    """Build a heatmap pool (n,17,64,48) fp32 directly in HBM, same recipe as `heatmaps`
    but with a per-track smooth drift instead of a sequential random walk so it vectorises.
    Returns (H, is_prev u8, is_next u8, boxes fp32 xyxy) as CUDA tensors."""
    import torch
    rng = np.random.default_rng(seed)
    ids, ip, inx = track_flags(n, rng, mean_len)
    g = torch.Generator(device=device).manual_seed(seed)
    H = torch.empty((n, J, HM_H, HM_W), dtype=torch.float32, device=device)
    ys = torch.arange(HM_H, device=device, dtype=torch.float32).view(1, 1, HM_H, 1)
    xs = torch.arange(HM_W, device=device, dtype=torch.float32).view(1, 1, 1, HM_W)
    ids_t = torch.from_numpy(ids).to(device)
    # per-track base centre and velocity; position inside the track drives the drift
    ntr = int(ids.max()) + 1
    base = torch.rand((ntr, J, 2), generator=g, device=device)
    vel = (torch.rand((ntr, J, 2), generator=g, device=device) - 0.5) * 2.0
    start = torch.from_numpy(np.r_[0, np.flatnonzero(np.diff(ids)) + 1]).to(device)
    pos_in = torch.arange(n, device=device) - start[ids_t]
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        t = ids_t[a:b]
        p = pos_in[a:b].view(-1, 1).float()
        cx = (6 + base[t, :, 0] * (HM_W - 12) + vel[t, :, 0] * p).clamp(2, HM_W - 3)
        cy = (6 + base[t, :, 1] * (HM_H - 12) + vel[t, :, 1] * p).clamp(2, HM_H - 3)
        amp = 0.3 + 0.7 * torch.rand((b - a, J), generator=g, device=device)
        blob = amp[:, :, None, None] * torch.exp(
            -((xs - cx[:, :, None, None]) ** 2 + (ys - cy[:, :, None, None]) ** 2) / 8.0)
        sec = (torch.rand((b - a, J), generator=g, device=device) < 0.1).float()
        sx = 3 + torch.rand((b - a, J), generator=g, device=device) * (HM_W - 7)
        sy = 3 + torch.rand((b - a, J), generator=g, device=device) * (HM_H - 7)
        sa = amp * (0.4 + 0.5 * torch.rand((b - a, J), generator=g, device=device)) * sec
        blob = blob + sa[:, :, None, None] * torch.exp(
            -((xs - sx[:, :, None, None]) ** 2 + (ys - sy[:, :, None, None]) ** 2) / 8.0)
        H[a:b] = blob + 0.02 * torch.randn(blob.shape, generator=g, device=device)
    bb = torch.from_numpy(boxes_xyxy(n, seed)).to(device)
    return H, torch.from_numpy(ip).to(device), torch.from_numpy(inx).to(device), bb


def device_embeddings(n: int, device, d: int = FEAT_D, seed: int = 2):
    """Clustered embeddings (see `embeddings`) built in HBM."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    nc = -(-n // 30)
    cen = torch.relu(torch.randn((nc, d), generator=g, device=device)) * 0.5
    X = torch.empty((n, d), dtype=torch.float32, device=device)
    step = 1 << 16
    for a in range(0, n, step):
        b = min(n, a + step)
        idx = torch.arange(a, b, device=device) // 30
        X[a:b] = cen[idx.clamp_max(nc - 1)] + 0.01 * torch.randn((b - a, d), generator=g, device=device)
    return X


# --------------------------------------------------------------------------------------------
# Counter-based pool ("hashed" generators): every value is a pure function of (seed, global
# item index, position), built from 32-bit integer hashes and single IEEE fp32 operations only,
# so numpy on the host and torch on the GPU produce the SAME BITS for any slice of the pool.
# That is what lets (a) every rank of an N-GPU run cut its shard out of ONE global pool,
# (b) the CPU reference run (which pins the goldens) see exactly the pool the GPU benchmark uses (tests/golden/coreset_scale.npz),
# (c) a pool larger than HBM be regenerated chunk by chunk.  Recipe and shapes follow SURVEY.md §8d
# (sigma-2 blobs on drifting centres, N(0,0.02)-like noise, geometric tracks of mean length 30,
# features = ceil(n/30) cluster centres ~ 0.5*ReLU(N(0,1)) + N(0,sigma) within-cluster noise);
# "normal" draws are Irwin-Hall sums of the four bytes of one hash word (mean 510, sd 147.8).
# --------------------------------------------------------------------------------------------
_M32 = 0xFFFFFFFF
_IH_MEAN, _IH_SD = 510, 147.8006
FEAT_KINDS = {"clustered": 0.01, "weak": 0.1, "iid": None}   # within-cluster sigma (30:1, 3:1 separation, none)
HEAT_RING = 250000       # heat-map content period of the global pool (see pool_heatmaps)


def _is_torch(a):
    return type(a).__module__.startswith("torch")


def _mix32(x):
    """32-bit integer hash evaluated in int64 (both multipliers < 2^31, so no product overflows
    int63): identical in numpy and torch."""
    x = x & _M32
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & _M32
    x = x ^ (x >> 15)
    x = (x * 0x2C1B3C6D) & _M32
    x = x ^ (x >> 16)
    return x


def _hash2(seed: int, stream: int, a, b):
    """hash of (seed, stream, a, b); a, b int64 arrays (broadcastable) with values < 2^32."""
    s = (int(seed) * 0x9E3779B1 + int(stream) * 0x85EBCA77 + 0x165667B1) & _M32
    return _mix32(_mix32(a ^ s) + b * 0x27D4EB2F)


def _ih4(h):
    """sum of the four bytes of a hash word minus its mean: integer in [-510, 510], ~N(0, 147.8^2)"""
    return (h & 255) + ((h >> 8) & 255) + ((h >> 16) & 255) + ((h >> 24) & 255) - _IH_MEAN


def _arange(lo, hi, device):
    if device is None:
        return np.arange(lo, hi, dtype=np.int64)
    import torch
    return torch.arange(lo, hi, dtype=torch.int64, device=device)


def _f32(k, scale_pow2: int):
    """float32(k) * 2^-scale_pow2 for an integer array |k| < 2^24: exact."""
    if _is_torch(k):
        import torch
        return k.to(torch.float32) * float(2.0 ** -scale_pow2)
    return k.astype(np.float32) * np.float32(2.0 ** -scale_pow2)


def pool_embeddings(n: int, lo: int = 0, hi: int | None = None, d: int = FEAT_D, seed: int = 2,
                    kind: str = "clustered", device=None, step: int = 1 << 15):
    """Rows [lo, hi) of the global (n, d) fp32 feature pool.  Values are K * 2^-20 with K an integer
    < 2^24 (exact in fp32): K = 3547*relu(g_c) + s*g_r with g the Irwin-Hall integers above, i.e.
    centre ~ 0.5*ReLU(N(0,1)) shared by rows i//30 == c, noise sigma = FEAT_KINDS[kind]
    ("iid": every row is its own ReLU(N(0,1)) draw).  device None -> numpy, else a CUDA torch tensor."""
    hi = n if hi is None else hi
    sig = FEAT_KINDS[kind]
    nscale = 0 if sig is None else int(round(sig / _IH_SD * (1 << 20)))
    if device is None:
        out = np.empty((hi - lo, d), dtype=np.float32)
    else:
        import torch
        out = torch.empty((hi - lo, d), dtype=torch.float32, device=device)
    cols = _arange(0, d, device)[None, :]
    for a in range(lo, hi, step):
        b = min(hi, a + step)
        rows = _arange(a, b, device)[:, None]
        if sig is None:
            k = _ih4(_hash2(seed, 3, rows, cols))
            k = (k * (k > 0)) * 7094
        else:
            g = _ih4(_hash2(seed, 1, rows // 30, cols))
            k = (g * (g > 0)) * 3547 + _ih4(_hash2(seed, 2, rows, cols)) * nscale
        out[a - lo:b - lo] = _f32(k, 20)
    return out


def pool_unc(n: int, lo: int = 0, hi: int | None = None, seed: int = 3, device=None):
    """U[0,1) float64 uncertainties of items [lo, hi) (24 random bits each, exact)."""
    hi = n if hi is None else hi
    i = _arange(lo, hi, device)
    h = _hash2(seed, 4, i, i * 0) >> 8
    if device is None:
        return h.astype(np.float64) * 2.0 ** -24
    import torch
    return h.to(torch.float64) * 2.0 ** -24


def pool_labeled(n: int, n_lab: int, seed: int = 5) -> np.ndarray:
    """The already-labelled subset (sorted global indices): the n_lab items with the smallest hash."""
    if n_lab <= 0:
        return np.zeros(0, dtype=np.int64)
    i = np.arange(n, dtype=np.int64)
    h = _hash2(seed, 5, i, i * 0) * (1 << 20) + (i & 0xFFFFF)     # (ties broken by index bits)
    return np.sort(np.argpartition(h, n_lab - 1)[:n_lab]).astype(np.int64)


def pool_tracks(n: int, seed: int = 0, mean_len: float = 30.0, ring: int = HEAT_RING):
    """Track structure of the global pool: geometric track lengths (numpy Generator, cheap, host)
    with a forced break at every multiple of `ring`.  Returns (track_id, pos_in_track, is_prev, is_next)
    as numpy arrays of length n."""
    ids, _, _ = track_flags(n, np.random.default_rng(seed), mean_len)
    brk = np.zeros(n, dtype=bool)
    brk[0] = True
    brk[1:] = ids[1:] != ids[:-1]
    if ring and ring < n:
        brk[np.arange(ring, n, ring)] = True
    tid = np.cumsum(brk) - 1
    start = np.flatnonzero(brk)
    pos = np.arange(n) - start[tid]
    is_prev = (~brk).astype(np.uint8)
    is_next = np.zeros(n, dtype=np.uint8)
    is_next[:-1] = is_prev[1:]
    return tid.astype(np.int64), pos.astype(np.int64), is_prev, is_next


_G_TABLE = None


def _gauss_table():
    """exp(-r^2/8) for r^2 = t/16, t = 0 .. (4*63)^2 + (4*47)^2 (quarter-pixel centres), float64 exp
    rounded once to fp32."""
    global _G_TABLE
    if _G_TABLE is None:
        t = np.arange((4 * (HM_H - 1)) ** 2 + (4 * (HM_W - 1)) ** 2 + 1, dtype=np.float64)
        _G_TABLE = np.exp(-t / 128.0).astype(np.float32)
    return _G_TABLE


def pool_heatmaps(track_id, pos_in_track, lo: int, hi: int, seed: int = 0, device=None, ring: int = HEAT_RING,
                  step: int = 1024, out=None):
    """Heat maps (hi-lo, 17, 64, 48) fp32 of pool items [lo, hi).  Item i's map is a function of
    (seed, i mod ring, its track id and position): a sigma-2 blob per joint whose centre (quarter-pixel
    grid) drifts linearly inside a track, amplitude in [0.3,1), with p ~ 0.1 a secondary blob of 0.4-0.9x
    amplitude, plus ~N(0,0.02) pixel noise.  The content repeats with period `ring` items: a 1 M-frame pool
    (208.9 GB of maps) does not fit one GPU, so the maps of items i and i+ring are equal (their boxes,
    features and scores are not); pass ring=0 for a pool without repetition."""
    m = hi - lo
    G = _gauss_table()
    if device is None:
        H = np.empty((m, J, HM_H, HM_W), dtype=np.float32) if out is None else out
        Gt = G
    else:
        import torch
        H = torch.empty((m, J, HM_H, HM_W), dtype=torch.float32, device=device) if out is None else out
        Gt = torch.from_numpy(G).to(device)
    tid_all, pos_all = np.asarray(track_id), np.asarray(pos_in_track)
    jj = _arange(0, J, device)[None, :]
    ys = (_arange(0, HM_H, device) * 4)[None, None, :, None]
    xs = (_arange(0, HM_W, device) * 4)[None, None, None, :]
    pix = (_arange(0, HM_H * HM_W, device)).reshape(1, 1, HM_H, HM_W)
    for a in range(lo, hi, step):
        b = min(hi, a + step)
        it = np.arange(a, b, dtype=np.int64)
        key = it % ring if ring else it
        t_np, p_np = tid_all[a:b], pos_all[a:b]
        if ring:   # the ring repeats the track structure too (pool_tracks breaks tracks at ring multiples)
            t_np, p_np = tid_all[key], pos_all[key]
        if device is None:
            t, p, key_ = t_np[:, None], p_np[:, None], key[:, None]
        else:
            import torch
            t = torch.from_numpy(t_np).to(device)[:, None]
            p = torch.from_numpy(p_np).to(device)[:, None]
            key_ = torch.from_numpy(key).to(device)[:, None]
        hb = _hash2(seed, 10, t, jj)                       # per (track, joint): base centre + velocity
        bx, by = hb & 127, (hb >> 7) & 127                 # quarter pixels over 32 px
        vx, vy = ((hb >> 14) & 7) - 3, ((hb >> 17) & 7) - 3    # quarter pixels per frame, |v| <= 0.75 px
        cx = 32 + bx + vx * p
        cy = 64 + by + vy * p
        cx = cx.clip(8, 4 * (HM_W - 3)) if device is None else cx.clamp(8, 4 * (HM_W - 3))
        cy = cy.clip(8, 4 * (HM_H - 3)) if device is None else cy.clamp(8, 4 * (HM_H - 3))
        hf = _hash2(seed, 11, key_, jj)                    # per (item, joint): amplitude, secondary blob
        amp = _f32(19661 + (((hf & 0xFFFF) * 45875) >> 16), 16)           # [0.3, 1.0)
        sec = ((hf >> 16) & 1023) < 102                                  # p = 0.0996
        hs = _hash2(seed, 12, key_, jj)
        sx = 12 + (hs & 127) + ((hs >> 7) & 31)            # quarter pixels, [3, 42.5] px
        sy = 12 + ((hs >> 12) & 255) - ((hs >> 20) & 31)   # [~-4.75+3, 66.75] -> clamped below
        sy = sy.clip(12, 4 * (HM_H - 4)) if device is None else sy.clamp(12, 4 * (HM_H - 4))
        samp = amp * _f32(26214 + ((((hs >> 25) & 127) * 258)), 16) * _f32(sec * 1, 0)   # 0.4 .. 0.9 x amp
        r1 = (xs - cx[:, :, None, None]) ** 2 + (ys - cy[:, :, None, None]) ** 2
        r2 = (xs - sx[:, :, None, None]) ** 2 + (ys - sy[:, :, None, None]) ** 2
        noise = _f32(_ih4(_hash2(seed, 13, (key_ * J + jj)[:, :, None, None], pix)) * 142, 20)   # sd 0.02
        blob = amp[:, :, None, None] * Gt[r1] + samp[:, :, None, None] * Gt[r2]
        H[a - lo:b - lo] = blob + noise
    return H


def pool_boxes(n: int, lo: int = 0, hi: int | None = None, seed: int = 0, device=None):
    """(hi-lo, 4) fp32 xyxy crop boxes of items [lo, hi): centre in [100,1000)x[100,700), height in
    [80,400), aspect 0.75; quarter-pixel integers, exact in fp32."""
    hi = n if hi is None else hi
    i = _arange(lo, hi, device)
    h1, h2 = _hash2(seed, 20, i, i * 0), _hash2(seed, 21, i, i * 0)
    cx = 400 + (h1 & 0xFFFF) * 3600 // 65536          # quarter pixels
    cy = 400 + (h1 >> 16) * 2400 // 65536
    hh = 320 + (h2 & 0xFFFF) * 1280 // 65536
    hh = hh - (hh % 8)                                 # height multiple of 2 px -> width 0.75 h exact
    ww = hh * 3 // 4
    cols = [cx - ww // 2, cy - hh // 2, cx - ww // 2 + ww, cy - hh // 2 + hh]
    if device is None:
        return _f32(np.stack(cols, axis=1), 2)
    import torch
    return _f32(torch.stack(cols, dim=1), 2)


def rank_pool(n: int, lo: int, hi: int, device, kind: str = "clustered", d: int = FEAT_D):
    """Items [lo, hi) of the global counter-based pool, resident on `device`: returns
    (heat-map segments [(pos, tensor)], boxes, is_prev, is_next, features, distinct frames held).
    The heat-map content has period HEAT_RING items (a 1 M-frame pool is 208.9 GB of maps, more than one
    GPU holds), so a range that spans whole periods keeps ONE copy of the ring and is scanned once per
    period — the same HBM traffic as distinct frames."""
    import torch
    ring = HEAT_RING
    tid, pos, ip, inx = pool_tracks(n, seed=0, ring=ring)
    cuts = [lo]
    while cuts[-1] < hi:
        cuts.append(min(hi, (cuts[-1] // ring + 1) * ring))
    segs = []
    if hi - lo >= ring:
        R = pool_heatmaps(tid, pos, 0, ring, seed=0, device=device, ring=ring)
        for a, b in zip(cuts[:-1], cuts[1:]):
            segs.append((a - lo, R[a % ring:a % ring + (b - a)]))
        distinct = ring
    else:
        for a, b in zip(cuts[:-1], cuts[1:]):
            segs.append((a - lo, pool_heatmaps(tid, pos, a, b, seed=0, device=device, ring=ring)))
        distinct = hi - lo
    bb = pool_boxes(n, lo, hi, seed=0, device=device)
    X = pool_embeddings(n, lo, hi, d=d, seed=2, kind=kind, device=device)
    ipd = torch.from_numpy(ip[lo:hi].copy()).to(device)
    inxd = torch.from_numpy(inx[lo:hi].copy()).to(device)
    return segs, bb, ipd, inxd, X, distinct
