"""TEST INFRASTRUCTURE — a NumPy restatement of scikit-learn's KMeans as the K-Means / weighted K-Means filters use it
(ActiveLearning.py:568-570, 596-598: `KMeans(n_clusters=query_size, random_state=318).fit_predict(embeddings[,
sample_weight])`), written step for step the way csrc/kmeans.cu + kmeans.py run it on the device.

The algorithm lives in a third-party dependency (scikit-learn 1.7.1 pinned by the reference, requirements.txt:174; 1.9.0
in this image): sklearn/cluster/_kmeans.py `_kmeans_plusplus` (:180-282), `_tolerance` (:285-293),
`_kmeans_single_lloyd` (:630-758), `KMeans.fit` (:1436-1565); sklearn/cluster/_k_means_lloyd.pyx
`lloyd_iter_chunked_dense` / `_update_chunk_dense`; sklearn/cluster/_k_means_common.pyx
`_relocate_empty_clusters_dense` (:167-211), `_average_centers` (:274-295), `_center_shift` (:298-311).

Pinned (tests/test_kmeans_golden_cpu.py): labels, iteration counts and queried rows equal sklearn's own on the
golden pools of tests/golden/kmeans.npz (written from the reference's statements) and on further pools, including
empty-cluster relocation and two-member clusters.  Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np


def seeds(X, w, k, rs):
    """_kmeans_plusplus: the first centre by `choice`, then per step 2 + int(log k) candidates drawn by searchsorted
    into the cumulative sum of w * closest_dist_sq, the candidate with the smallest potential wins.  The k-1 later
    uniform(size=trials) draws are taken as ONE block (same stream)."""
    n = X.shape[0]
    first = int(rs.choice(n, p=w / w.sum()))
    trials = 2 + int(np.log(k))
    rand = rs.uniform(size=max(k - 1, 0) * trials).reshape(max(k - 1, 0), trials)
    xx = (X * X).sum(1)
    ids = [first]
    closest = np.maximum((-2 * (X @ X[first]) + xx[first]) + xx, 0)
    pot = (closest * w).sum()
    for c in range(1, k):
        cand = np.minimum(np.searchsorted(np.cumsum(w * closest), rand[c - 1] * pot), n - 1)
        D = np.maximum((-2 * (X[cand] @ X.T) + xx[cand][:, None]) + xx[None, :], 0)
        D = np.minimum(D, closest[None, :])
        pots = D @ w
        b = int(np.argmin(pots))
        pot, closest = pots[b], D[b]
        ids.append(int(cand[b]))
    return np.asarray(ids)


def lloyd(X, w, seed_ids, tol_abs, max_iter=300):
    """_kmeans_single_lloyd on centres kept in the centred frame (KMeans.fit subtracts X.mean(axis=0)); the E step runs
    in the raw frame on centre + mean (label decisions are robust to the translation), the M step sums
    (x - mean) * w in ascending row order.  Returns labels, raw-frame centres, iterations, relocations."""
    n, k = X.shape[0], len(seed_ids)
    mean = X.mean(axis=0)
    Xc = X - mean
    C = X[seed_ids] - mean
    labels_old = np.full(n, -1)
    strict, reloc = False, 0
    for it in range(max_iter):
        Cr = C + mean
        labels = ((Cr * Cr).sum(1)[None, :] - 2 * X @ Cr.T).argmin(1)
        sums, wsum = np.zeros_like(C), np.zeros(k)
        np.add.at(sums, labels, Xc * w[:, None])
        np.add.at(wsum, labels, w)
        empty = np.where(wsum == 0)[0]
        if len(empty):                                       # _relocate_empty_clusters_dense
            dist = ((X - Cr[labels]) ** 2).sum(1)
            far = np.argpartition(dist, -len(empty))[:-len(empty) - 1:-1]
            if dist.max() != 0:
                for e, f in zip(empty, far):
                    o = labels[f]
                    sums[o] -= Xc[f] * w[f]
                    sums[e] = Xc[f] * w[f]
                    wsum[e] = w[f]
                    wsum[o] -= w[f]
                    reloc += 1
        Cn, am = np.empty_like(C), int(np.argmax(wsum))      # _average_centers (empty: the biggest cluster's row as it stands)
        for j in range(k):
            Cn[j] = sums[j] * (1.0 / wsum[j]) if wsum[j] > 0 else (Cn[am] if am < j else sums[am])
        shift = np.sqrt(((Cn - C) ** 2).sum(1))              # _center_shift
        C = Cn
        if np.array_equal(labels, labels_old):
            strict = True
            break
        if (shift ** 2).sum() <= tol_abs:
            break
        labels_old = labels
    Cr = C + mean
    if not strict:
        labels = ((Cr * Cr).sum(1)[None, :] - 2 * X @ Cr.T).argmin(1)
    return labels, Cr, it + 1, reloc


def fit_select(X, k, w=None, seed=318, tol=1e-4):
    """fit_predict + the reference's per-cluster closest member (:599-603).  X float array (n,d)."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    w = np.ones(n) if w is None else np.asarray(w, dtype=np.float64)
    rs = np.random.RandomState(seed)
    ids = seeds(X, w, k, rs)
    labels, centers, n_iter, reloc = lloyd(X, w, ids, np.var(X, axis=0).mean() * tol)
    dis = ((X - centers[labels]) ** 2).sum(axis=1)
    cluster_num = len(np.unique(labels))
    rows = [int(np.arange(n)[labels == i][dis[labels == i].argmin()]) for i in range(cluster_num)]
    return rows, labels, n_iter, reloc, ids
