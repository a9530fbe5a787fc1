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
