"""CPU ORACLE for the VATL4Pose active-learning query pass.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the arithmetic of the reference's query pass
(ImIntheMiddle/VATL4Pose-WACV2024, `active_learning/ActiveLearning.py:253-649`) with the
same third-party numerics the reference calls (numpy reductions, scipy.ndimage
maximum_filter, sklearn pairwise_distances, cv2.getAffineTransform, torch nn.Linear).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it, and only as the checker / the timed CPU arm.  The product package
never imports it and has no CPU fallback.

Pinning: every function below is compared BIT-FOR-BIT against the reference's own function
(imported from /root/reference with import stubs) by `oracle/pin_against_reference.py`,
which also writes the golden fixtures under `tests/golden/`.  The reference ships no tests
or golden vectors for this path, so the pin is "outputs of the reference
itself run in the build container".  At BASELINE scale (100 000 .. 1 000 000 items) the selection is
pinned by `oracle/pin_scale.py`, which runs the reference's own `coreset_selection` on the pools the
benchmark uses (`tests/golden/coreset_scale_*.npz`).
ONE EXCEPTION — PARITY UNPINNED: `peak_local_max` below restates scikit-image 0.24's published algorithm
(scikit-image is not installed in the build container); MPE / Margin are pinned through the reference's own
`compute_mpe` / `compute_margin` calling that restatement, not against the library itself.

Citations are `path:line` relative to the reference root.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# a-1  temporal heatmap continuity
# --------------------------------------------------------------------------------------

def thc_pair(cur: np.ndarray, adj: np.ndarray, norm_type: str = "L1"):
    """active_learning/ActiveLearning.py:747-760 — sum over all J*H*W of |cur-adj|
    (or squared), divided by the number of joints.  fp32 numpy pairwise sum."""
    n_joints = cur.shape[0]
    delta = cur - adj
    if norm_type == "L1":
        return np.sum(np.abs(delta)) / n_joints
    if norm_type == "L2":
        return np.sum(np.square(delta)) / n_joints
    raise ValueError(norm_type)


def thc_item(cur, prev, nxt, has_prev: bool, has_next: bool) -> float:
    """Call-site logic active_learning/ActiveLearning.py:345-363: add the pair terms that
    exist, double when exactly one neighbour exists, 0 when none."""
    acc = 0
    if has_prev:
        acc += thc_pair(cur, prev)
    if has_next:
        acc += thc_pair(cur, nxt)
        if not has_prev:
            acc *= 2
    elif has_prev:
        acc *= 2
    return float(acc)


def thc_pool(H: np.ndarray, is_prev, is_next, halo_prev=None, halo_next=None) -> np.ndarray:
    """THC of every item of an id-sorted pool, neighbours taken from the pool itself
    (pred_prev[i] == pred[i-1] whenever isPrev[i]).  float64 vector, each
    entry the `float(thc)` of ActiveLearning.py:363."""
    n = H.shape[0]
    out = np.zeros(n, dtype=np.float64)
    for i in range(n):
        p = H[i - 1] if i > 0 else halo_prev
        q = H[i + 1] if i < n - 1 else halo_next
        hp = bool(is_prev[i]) and p is not None
        hn = bool(is_next[i]) and q is not None
        out[i] = thc_item(H[i], p, q, hp, hn)
    return out


# --------------------------------------------------------------------------------------
# a-2  local peaks
# --------------------------------------------------------------------------------------

def localpeak_values(img: np.ndarray, filter_size: int = 3, order: float = 0.5) -> np.ndarray:
    """active_learning/local_peak.py:5-10 — pixels equal to their zero-padded
    filter_size x filter_size maximum, kept when >= order * (largest such pixel)."""
    from scipy.ndimage import maximum_filter
    win = maximum_filter(img, footprint=np.ones((filter_size, filter_size)), mode="constant")
    is_peak = img == win
    if not is_peak.any():
        return img[is_peak]  # empty, same dtype
    top = img[is_peak].max()
    keep = is_peak & (img >= top * order)
    return img[keep]  # row-major order == masked-array .compressed()


def localpeak_mean(hm: np.ndarray, filter_size: int = 3, order: float = 0.5):
    """active_learning/local_peak.py:12-22 — mean of the kept peak values of all joints
    (NaN with a RuntimeWarning if nothing survives)."""
    vals = [localpeak_values(m, filter_size, order) for m in hm]
    return np.hstack(vals).mean()


# --------------------------------------------------------------------------------------
# a-3  heatmap -> coordinates
# --------------------------------------------------------------------------------------

def max_pred(hm: np.ndarray):
    """alphapose/utils/transforms.py:710-727 — first flat argmax per joint as (x,y) fp32,
    zeroed when the maximum is <= 0; maxvals (J,1)."""
    nj, _, w = hm.shape
    flat = hm.reshape(nj, -1)
    vmax = flat.max(axis=1).reshape(nj, 1)
    where = flat.argmax(axis=1).reshape(nj, 1)
    xy = np.tile(where, (1, 2)).astype(np.float32)
    xy[:, 0] = xy[:, 0] % w
    xy[:, 1] = np.floor(xy[:, 1] / w)
    xy *= np.tile(vmax > 0.0, (1, 2)).astype(np.float32)
    return xy, vmax


def heatmap_coords(hm: np.ndarray):
    """Heat-map-space part of alphapose/utils/transforms.py:550-566: argmax plus the
    quarter-pixel shift toward the larger neighbour, strictly interior pixels only."""
    xy, vmax = max_pred(hm)
    h, w = hm.shape[1], hm.shape[2]
    for j in range(xy.shape[0]):
        px = int(round(float(xy[j][0])))
        py = int(round(float(xy[j][1])))
        if 1 < px < w - 1 and 1 < py < h - 1:
            d = np.array((hm[j][py][px + 1] - hm[j][py][px - 1],
                          hm[j][py + 1][px] - hm[j][py - 1][px]))
            xy[j] += np.sign(d) * .25
    return xy, vmax


def _third_point(a, b):
    d = a - b
    return b + np.array([-d[1], d[0]], dtype=np.float32)


def inverse_affine(center, scale, out_size):
    """alphapose/utils/transforms.py:753-786 with rot=0, inv=1: three float32 point pairs,
    solved by cv2.getAffineTransform in double.  Only scale[0] (the width) is used."""
    import cv2
    src_w = scale[0]
    dst_w, dst_h = out_size
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0] = center
    src[1] = center + np.array([0.0, src_w * -0.5])  # get_dir with rot_rad = 0 (:795-803)
    dst[0] = [dst_w * 0.5, dst_h * 0.5]
    dst[1] = np.array([dst_w * 0.5, dst_h * 0.5]) + np.array([0, dst_w * -0.5], np.float32)
    src[2] = _third_point(src[0], src[1])
    dst[2] = _third_point(dst[0], dst[1])
    return cv2.getAffineTransform(np.float32(dst), np.float32(src))


def heatmap_to_coord(hm: np.ndarray, bbox_xyxy):
    """alphapose/utils/transforms.py:550-583 — image-space joint coordinates (J,2) fp32 and
    peak values (J,1) fp32 from one (J,H,W) heat-map stack and its crop box."""
    xy, vmax = heatmap_coords(hm)
    h, w = hm.shape[1], hm.shape[2]
    xmin, ymin, xmax, ymax = bbox_xyxy
    bw = xmax - xmin
    bh = ymax - ymin
    center = np.array([xmin + bw * 0.5, ymin + bh * 0.5])
    scale = np.array([bw, bh])
    out = np.zeros_like(xy)
    for j in range(xy.shape[0]):
        t = inverse_affine(center, scale, [w, h])
        out[j] = np.dot(t, np.array([xy[j][0], xy[j][1], 1.]).T)[:2]
    return out, vmax


def keypoints_row(coords: np.ndarray, vmax: np.ndarray):
    """active_learning/ActiveLearning.py:306-307 — [x0,y0,s0,x1,...] python floats."""
    return np.concatenate((coords, vmax), axis=1).reshape(-1).tolist()


# --------------------------------------------------------------------------------------
# a-4 / a-5  whole-body pose unnaturalness
# --------------------------------------------------------------------------------------

def xyxy_to_xywh(b):
    """alphapose/utils/bbox.py:91-97 (tuple/list branch): w = x2-x1+1, h = y2-y1+1."""
    return (b[0], b[1], b[2] - b[0] + 1, b[3] - b[1] + 1)


TRIANGLES = ((8, 6, 12), (6, 8, 10), (5, 7, 9), (7, 5, 11),
             (11, 12, 14), (12, 11, 13), (12, 14, 16), (11, 13, 15))


def _angle(x0, y0, x1, y1, x2, y2):
    """active_learning/Whole_body_AE/hybrid_feature.py:6-12."""
    eps = 1e-6
    m1 = (y1 - y0) / (x1 - x0 + eps)
    m2 = (y2 - y1) / (x2 - x1 + eps)
    return np.arctan(np.abs((m1 - m2) / (1 + m1 * m2 + eps)))


def hybrid_feature(bbox_xywh, keypoints) -> np.ndarray:
    """active_learning/Whole_body_AE/hybrid_feature.py:14-58 — 42-d float64 feature:
    score-weighted-centroid-relative x and y over box height, then 8 limb angles."""
    height = bbox_xywh[3]
    assert height > 0, "height of human body must be positive!"
    xs = keypoints[0::3]
    ys = keypoints[1::3]
    sc = keypoints[2::3]
    assert sum(sc) > 0, "at least one visible keypoint is required!"
    gx = np.average(xs, weights=sc)
    gy = np.average(ys, weights=sc)
    fx = (np.array(xs) - gx) / height
    fy = (np.array(ys) - gy) / height
    ang = np.zeros(8)
    for t, (a, b, c) in enumerate(TRIANGLES):
        ang[t] = _angle(xs[a], ys[a], xs[b], ys[b], xs[c], ys[c])
    return np.hstack((fx, fy, ang))


def make_autoencoder(weights):
    """torch module with the layer stack of active_learning/Whole_body_AE/AutoEncoder.py:13-32,
    dimensions taken from `weights` (list of 8 (W[out,in], b[out]) fp32 pairs)."""
    import torch
    import torch.nn as nn
    layers = []
    for k, (W, b) in enumerate(weights):
        lin = nn.Linear(W.shape[1], W.shape[0])
        with torch.no_grad():
            lin.weight.copy_(torch.from_numpy(np.ascontiguousarray(W)))
            lin.bias.copy_(torch.from_numpy(np.ascontiguousarray(b)))
        layers.append(lin)
        if k == 3:
            continue  # no activation between encoder output and decoder input
        layers.append(nn.Sigmoid() if k == 7 else nn.ReLU(True))
    return nn.Sequential(*layers).eval()


def wpu_item(ae, bbox_xyxy, keypoints, drop_ears: bool = False) -> float:
    """active_learning/ActiveLearning.py:364-370 (THC+WPU: MSE over all dims) and :371-386
    (WPU only: ear dims {3,4,20,21} removed before the MSE).  `ae` from make_autoencoder."""
    import torch
    feat = hybrid_feature(xyxy_to_xywh(bbox_xyxy), keypoints)
    u = torch.tensor(feat).float()
    with torch.no_grad():
        r = ae(u)
    if drop_ears:
        a, b = u.numpy(), r.numpy()
        a = np.concatenate([a[:3], a[5:20], a[22:]])
        b = np.concatenate([b[:3], b[5:20], b[22:]])
        u, r = torch.tensor(a).float(), torch.tensor(b).float()
    return float(torch.nn.functional.mse_loss(r, u))


# --------------------------------------------------------------------------------------
# a-6  score fusion
# --------------------------------------------------------------------------------------

def _minmax(v):
    return (v - np.min(v)) / (np.max(v) - np.min(v))


def fuse_scores(thc_u, wpu_u=None, mode: str = "const", labeled_ratio: float = 0.0) -> np.ndarray:
    """active_learning/ActiveLearning.py:490-516 over the unlabelled items (ascending index):
    per-criterion min-max, combine (const / increase / decrease), min-max again.  With a
    single criterion (wpu_u None) just one min-max (:511-516).  |U| <= 1 -> zeros (:490)."""
    thc_u = np.asarray(thc_u, dtype=np.float64)
    if thc_u.size <= 1:
        return np.zeros(thc_u.size)
    if wpu_u is None:
        return _minmax(thc_u)
    a = _minmax(thc_u)
    b = _minmax(np.asarray(wpu_u, dtype=np.float64))
    if mode == "const":
        u = a + b
    elif mode == "increase":
        u = labeled_ratio * a + (1 - labeled_ratio) * b
    elif mode == "decrease":
        u = (1 - labeled_ratio) * a + labeled_ratio * b
    else:
        raise ValueError(mode)
    return _minmax(u)


# --------------------------------------------------------------------------------------
# a-7  k-center greedy core-set
# --------------------------------------------------------------------------------------

def coreset_select(X: np.ndarray, unc: np.ndarray, labeled, k: int, moks: float, lam: float,
                   rule: str = "w_unc", first_pick: int | None = None):
    """active_learning/ActiveLearning.py:798-850.  X float64 (N,D); unc float64 (N,), modified
    in place like the reference; labeled: iterable of already-labelled indices.
    rule: "w_unc" (:815-821), "fixed_lambda" (:822-827) or "dist" (:828-833; the random first
    pick of an empty labelled set is supplied by the caller as `first_pick`)."""
    from sklearn.metrics import pairwise_distances
    lab = np.asarray(list(labeled), dtype=np.int64)
    n_lab = lab.size
    picks = []
    if n_lab:
        md = pairwise_distances(X, X[lab], metric="euclidean").min(axis=1).reshape(-1, 1)
    else:
        md = None
    for _ in range(k):
        if n_lab == 0:
            if rule == "dist":
                ind = int(first_pick)
            else:
                ind = np.argmax(unc)
        elif rule == "w_unc":
            ind = np.argmax(((1 - moks) * md.reshape(-1)) + (lam * moks * unc))
        elif rule == "fixed_lambda":
            ind = np.argmax(md.reshape(-1) + lam * unc)
        else:
            ind = np.argmax(md.reshape(-1))
        col = pairwise_distances(X, X[[ind]], metric="euclidean")
        md = col.min(axis=1).reshape(-1, 1) if md is None else np.minimum(md, col)
        n_lab += 1
        unc[ind] = 0
        picks.append(int(ind))
    return picks, (None if md is None else md.reshape(-1))


# --------------------------------------------------------------------------------------
# SURVEY.md §8f "next" rows: HP / TPC / Entropy uncertainties, Influence, Diversity, top-k
# --------------------------------------------------------------------------------------

def hp_item(pose_scores) -> float:
    """active_learning/ActiveLearning.py:329-330 — highest-probability uncertainty: minus the
    sum of the 17 peak values (fp32 numpy sum over the (17,1) maxvals array)."""
    return float(-np.sum(pose_scores))


def tpc_pair(cur_pose: np.ndarray, adj_hm: np.ndarray, bbox_xyxy, thresh) -> int:
    """active_learning/ActiveLearning.py:736-745 (compute_tpc) — joints whose image-space
    position differs by more than `thresh` between the current pose and the pose decoded from
    the adjacent frame's heat maps with the CURRENT crop box."""
    adj_pose, _ = heatmap_to_coord(adj_hm, bbox_xyxy)
    dist = np.linalg.norm(cur_pose - adj_pose, axis=1)
    return np.count_nonzero(dist > thresh)


def tpc_item(cur_pose, prev_hm, next_hm, bbox_xyxy, has_prev: bool, has_next: bool) -> float:
    """Call-site logic active_learning/ActiveLearning.py:333-344."""
    thresh = 0.01 * np.sqrt((bbox_xyxy[2] - bbox_xyxy[0]) * (bbox_xyxy[3] - bbox_xyxy[1]))
    tpc = 0
    if has_prev:
        tpc += tpc_pair(cur_pose, prev_hm, bbox_xyxy, thresh)
    if has_next:
        tpc += tpc_pair(cur_pose, next_hm, bbox_xyxy, thresh)
        if not has_prev:
            tpc *= 2
    elif has_prev:
        tpc *= 2
    return float(tpc)


def entropy_item(hm: np.ndarray) -> float:
    """active_learning/ActiveLearning.py:790-796 (compute_entropy) — sum over the joints of
    scipy.stats.entropy of the flattened (raw, un-normalised) map."""
    from scipy.stats import entropy
    total = 0
    for m in hm:
        total += entropy(m.flatten())
    return float(total)


def pose_unc_pool(H, boxes_xyxy, is_prev, is_next, kind: str) -> np.ndarray:
    """HP / TPC / Entropy of every item of an id-sorted pool (neighbours from the pool itself,
    like thc_pool).  float64 vector of the `float(uncertainty)` values."""
    n = H.shape[0]
    out = np.zeros(n, dtype=np.float64)
    for i in range(n):
        box = [float(v) for v in boxes_xyxy[i]]
        if kind == "Entropy":
            out[i] = entropy_item(H[i])
            continue
        if kind in ("MPE", "Margin"):
            out[i] = mpe_item(H[i]) if kind == "MPE" else margin_item(H[i])
            continue
        pose, scores = heatmap_to_coord(H[i], box)
        if kind == "HP":
            out[i] = hp_item(scores)
        elif kind == "TPC":
            hp_ = bool(is_prev[i]) and i > 0
            hn_ = bool(is_next[i]) and i < n - 1
            out[i] = tpc_item(pose, H[i - 1] if hp_ else None, H[i + 1] if hn_ else None, box, hp_, hn_)
        else:
            raise ValueError(kind)
    return out


def cosine_rowsum(Xs: np.ndarray) -> np.ndarray:
    """active_learning/ActiveLearning.py:471-475 and :582-585 — row sums of the distance graph
    KNeighborsTransformer(mode='distance', metric='cosine', n_neighbors=m-1) builds over Xs."""
    from sklearn.neighbors import KNeighborsTransformer
    knn = KNeighborsTransformer(mode="distance", metric="cosine", n_neighbors=len(Xs) - 1)
    graph = knn.fit_transform(Xs)
    return np.asarray(np.sum(graph, axis=1)).flatten()


def influence_scores(X: np.ndarray, unlabeled_idx) -> np.ndarray:
    """active_learning/ActiveLearning.py:468-477 — min-max normalised cosine row sums over the
    unlabelled items; zeros when |U| <= 1."""
    if len(unlabeled_idx) in (0, 1):
        return np.zeros(len(unlabeled_idx))
    s = cosine_rowsum(X[unlabeled_idx])
    return _minmax(s)


def total_score(unc_score, influence, combine_weight):
    """active_learning/ActiveLearning.py:517-526."""
    if unc_score is not None and influence is not None:
        return combine_weight * unc_score + (1 - combine_weight) * influence
    return unc_score if unc_score is not None else influence


def candidate_order(unlabeled_idx, total):
    """active_learning/ActiveLearning.py:527-530 — unlabelled ids by descending score (stable)."""
    d = dict((idx, sc) for idx, sc in zip(unlabeled_idx, total))
    return [int(i) for i, _ in sorted(d.items(), key=lambda x: x[1], reverse=True)]


def topk_select(unlabeled_idx, total, k: int):
    """filter == "None" (:533-534,541-542): the k best ids, returned sorted ascending."""
    return sorted(candidate_order(unlabeled_idx, total)[:k])


def diversity_select(X: np.ndarray, unlabeled_idx, total, k: int):
    """filter == "Diversity" (:537-538,581-590): the 8k best ids (sorted ascending) re-ranked by
    ascending cosine row sum among themselves; the first k."""
    cand = sorted(candidate_order(unlabeled_idx, total)[:8 * k])
    div = cosine_rowsum(X[cand])
    d = dict((idx, sc) for idx, sc in zip(cand, div))
    return [int(i) for i, _ in sorted(d.items(), key=lambda x: x[1])][:k]


# --------------------------------------------------------------------------------------
# MPE / Margin (skimage.feature.peak_local_max based)
# --------------------------------------------------------------------------------------

def peak_local_max(image: np.ndarray, min_distance: int = 1, num_peaks=np.inf) -> np.ndarray:
    """RESTATEMENT of skimage.feature.peak_local_max (scikit-image 0.24.0, the version the reference pins in
    requirements.txt:172 / uv.lock; skimage/feature/peak.py) for the arguments the reference passes
    (ActiveLearning.py:773,784: min_distance=5, num_peaks=5; everything else default: threshold_abs=None,
    threshold_rel=None, exclude_border=True, p_norm=inf, no labels / footprint).  scikit-image is NOT installed
    in the build container, so this restatement follows the published algorithm and its parity with the library
    itself is UNPINNED:
      footprint = ones((2*min_distance+1,)*2); image_max = ndi.maximum_filter(image, footprint, mode='nearest');
      mask = image == image_max, all False when every pixel equals its window maximum (trivial image);
      mask &= image > threshold with threshold = image.min(); the border of width min_distance is excluded;
      candidates in np.nonzero order are stably sorted by descending intensity; ensure_spacing keeps a peak unless an
      already kept one lies at Chebyshev distance < min_distance; at most num_peaks are returned, (row, col)."""
    from scipy import ndimage as ndi
    size = 2 * min_distance + 1
    image_max = ndi.maximum_filter(image, footprint=np.ones((size, size), dtype=bool), mode="nearest")
    out = image == image_max
    if np.all(out):
        out[:] = False
    out &= image > image.min()
    b = min_distance
    if b > 0:
        out[:b, :] = False; out[-b:, :] = False; out[:, :b] = False; out[:, -b:] = False
    coord = np.nonzero(out)
    inten = image[coord]
    order = np.argsort(-inten, kind="stable")
    coord = np.transpose(coord)[order]
    kept = []
    max_out = int(num_peaks) if np.isfinite(num_peaks) else None
    for c in coord:
        if all(max(abs(int(c[0]) - int(k[0])), abs(int(c[1]) - int(k[1]))) >= min_distance for k in kept):
            kept.append(c)
            if max_out is not None and len(kept) >= max_out:
                break
    return np.array(kept, dtype=np.int64).reshape(-1, 2)


def mpe_item(hm: np.ndarray, plm=peak_local_max) -> float:
    """active_learning/ActiveLearning.py:762-778 (compute_mpe): entropy of the softmax of the <= 5 local peak
    values of every joint map, summed over the joints."""
    from scipy.special import softmax
    from scipy.stats import entropy
    mpe = 0
    for heatmap in hm:
        loc = plm(heatmap, min_distance=5, num_peaks=5)
        peaks = heatmap[loc[:, 0], loc[:, 1]]
        if peaks.shape[0] > 0:
            peaks = softmax(peaks)
            mpe += entropy(peaks)
    return float(mpe)


def margin_item(hm: np.ndarray, plm=peak_local_max) -> float:
    """active_learning/ActiveLearning.py:780-788 (compute_margin): |top peak - second peak| summed over the joints."""
    margin = 0
    for heatmap in hm:
        loc = plm(heatmap, min_distance=5, num_peaks=5)
        peaks = heatmap[loc[:, 0], loc[:, 1]]
        if peaks.shape[0] > 1:
            margin += np.linalg.norm(peaks[0] - peaks[1])
    return float(margin)


# --------------------------------------------------------------------------------------
# OKS (the evaluation the selection weights are derived from)
# --------------------------------------------------------------------------------------
OKS_SIGMAS = np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07, .87, .87, .89, .89]) / 10.0
OKS_VARS = (OKS_SIGMAS * 2) ** 2


def compute_oks(bb_xywh, predkpts, gtkpts) -> float:
    """active_learning/al_metric.py:42-69 — OKS of a predicted pose (51 values) against the ground
    truth, body area = GT box w*h; keypoints invisible in the GT are ignored, and when none is
    visible the distance to the doubled box is used."""
    d, g = np.array(predkpts), np.array(gtkpts)
    xg, yg, vg = g[0::3], g[1::3], g[2::3]
    k1 = np.count_nonzero(vg > 0)
    x0, x1 = bb_xywh[0] - bb_xywh[2], bb_xywh[0] + bb_xywh[2] * 2
    y0, y1 = bb_xywh[1] - bb_xywh[3], bb_xywh[1] + bb_xywh[3] * 2
    area = bb_xywh[2] * bb_xywh[3]
    xd, yd = d[0::3], d[1::3]
    if k1 > 0:
        dx, dy = xd - xg, yd - yg
    else:
        z = np.zeros(17)
        dx = np.max((z, x0 - xd), axis=0) + np.max((z, xd - x1), axis=0)
        dy = np.max((z, y0 - yd), axis=0) + np.max((z, yd - y1), axis=0)
    e = (dx ** 2 + dy ** 2) / OKS_VARS / (area + np.spacing(1)) * 0.5
    if k1 > 0:
        e = e[vg > 0]
    return np.sum(np.exp(-e)) / e.shape[0]


def mean_oks_of_queries(query_list, oks_by_idx) -> float:
    """active_learning/ActiveLearning.py:852-858 — mOKS of the newly queried items (in descending OKS
    order, as get_retrain_id sorts them before np.mean)."""
    vals = sorted((oks_by_idx[i] for i in query_list), reverse=True)
    return np.mean(vals)


# --------------------------------------------------------------------------------------
# whole scoring loop (the CPU arm that bench.py times)
# --------------------------------------------------------------------------------------

def score_pool(H, boxes_xyxy, is_prev, is_next, ae, unlabeled_mask=None, drop_ears=False,
               want=("coords", "thc", "wpu", "peak")):
    """Per-person loop of active_learning/ActiveLearning.py:299-414 restricted to the
    scoring arithmetic (no json / OKS / mAP).  Returns a dict of per-item arrays."""
    n = H.shape[0]
    out = {"kpts": np.zeros((n, 51)), "thc": np.zeros(n), "wpu": np.zeros(n),
           "peak": np.full(n, np.nan)}
    for i in range(n):
        box = [float(v) for v in boxes_xyxy[i]]
        if "coords" in want or "wpu" in want:
            c, v = heatmap_to_coord(H[i], box)
            kp = keypoints_row(c, v)
            out["kpts"][i] = kp
        if "thc" in want:
            p = H[i - 1] if i > 0 else None
            q = H[i + 1] if i < n - 1 else None
            out["thc"][i] = thc_item(H[i], p, q, bool(is_prev[i]) and p is not None,
                                     bool(is_next[i]) and q is not None)
        if "wpu" in want:
            out["wpu"][i] = wpu_item(ae, box, np.array(kp) if not drop_ears else kp, drop_ears)
        if "peak" in want and (unlabeled_mask is None or unlabeled_mask[i]):
            out["peak"][i] = localpeak_mean(H[i])
    return out


# --------------------------------------------------------------------------------------
# K-Means / weighted K-Means filters (ActiveLearning.py:593-608, :553-580).  The clustering itself is the
# third-party dependency the reference calls (scikit-learn, 1.7.1 pinned by the reference, 1.9.0 here): the oracle
# calls it too; the lines around it are restated.  pin_against_reference.py executes the reference's own source
# lines of both branches and compares.
# --------------------------------------------------------------------------------------
def _closest_member_per_cluster(embeddings, cluster_learner, cluster_idxs):
    cluster_num = len(np.unique(cluster_idxs))                                         # (:571,598)
    centers = cluster_learner.cluster_centers_[cluster_idxs]                           # (:572,599)
    dis = ((embeddings - centers) ** 2).sum(axis=1)                                    # (:573-574,600-601)
    return [int(np.arange(embeddings.shape[0])[cluster_idxs == i][dis[cluster_idxs == i].argmin()])
            for i in range(cluster_num)]                                               # (:575,602)


def kmeans_filter(fvecs_matrix, candidate_list, query_size, n_unlabeled=None):
    """filter == "K-Means" (:593-608): returns (query_list, query_size, labels)."""
    from sklearn.cluster import KMeans
    embeddings = np.asarray(fvecs_matrix, dtype=np.float64)[candidate_list]
    n_unlabeled = len(candidate_list) if n_unlabeled is None else n_unlabeled
    if n_unlabeled < query_size:
        query_size = n_unlabeled
    cluster_learner = KMeans(n_clusters=query_size, random_state=318)
    cluster_idxs = cluster_learner.fit_predict(embeddings)
    rows = _closest_member_per_cluster(embeddings, cluster_learner, cluster_idxs)
    return [int(candidate_list[i]) for i in rows], query_size, cluster_idxs


def weighted_kmeans_filter(fvecs_matrix, candidate_list, total_score, w_unc, combine_weight, query_size, n_unlabeled=None):
    """filter == "weighted" (:553-580): duplicates removed with np.unique(axis=0) (which also SORTS the rows), weights
    1 + w_unc * combine_weight * total_score; the reference maps the chosen rows of the de-duplicated, sorted matrix
    straight through candidate_list (:580) — kept as it is.  Returns (query_list, query_size, labels, embed_idx)."""
    from sklearn.cluster import KMeans
    embeddings = np.asarray(fvecs_matrix, dtype=np.float64)[candidate_list]
    _, embed_idx = np.unique(embeddings, axis=0, return_index=True)
    embeddings = embeddings[embed_idx]
    weight = 1 + w_unc * combine_weight * np.asarray(total_score, dtype=np.float64)
    weight = weight[embed_idx]
    n_unlabeled = len(candidate_list) if n_unlabeled is None else n_unlabeled
    if n_unlabeled <= query_size:
        query_size = n_unlabeled
    if query_size > len(embeddings):
        query_size = len(embeddings)
    cluster_learner = KMeans(n_clusters=query_size, random_state=318, verbose=0)
    cluster_idxs = cluster_learner.fit_predict(embeddings, sample_weight=weight)
    rows = _closest_member_per_cluster(embeddings, cluster_learner, cluster_idxs)
    return [int(candidate_list[i]) for i in rows], query_size, cluster_idxs, embed_idx
