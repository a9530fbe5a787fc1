"""K-Means / weighted K-Means filters on the device (csrc/kmeans.cu, kmeans.py) against the reference's own
statements (ActiveLearning.py:553-580, 593-608; golden written by oracle/pin_against_reference.py kmeans) and against
sklearn directly on further pools.  Bars: labels, query lists, iteration counts and de-duplication indices exact;
centres within 1e-9 (sklearn's own centre sums depend on its OpenMP chunking)."""
import os
import sys
import warnings
from types import SimpleNamespace

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kmeans_cases import CASES, case_inputs     # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", CASES)
def test_kmeans_filters_match_reference_golden(built_lib, tag):
    v = built_lib
    from vatlq import kmeans as KM
    z = np.load(os.path.join(ROOT, "tests", "golden", "kmeans.npz"))
    X, cand, score, k, w_unc, cw = case_inputs(z, tag)
    dev = "cuda:0"
    Xd = torch.from_numpy(X).to(dev)
    cand_t = torch.as_tensor(cand, device=dev)
    # K-Means
    res = KM.kmeans_fit_select(Xd[cand_t], k)
    assert np.array_equal(res.labels.cpu().numpy(), z[f"{tag}_km_labels"])
    assert [cand[i] for i in res.query_rows] == z[f"{tag}_km_query"].tolist()
    assert res.n_iter == int(z[f"{tag}_km_niter"])
    # weighted: np.unique(axis=0) order, weights, clamped query size
    eidx = KM.unique_rows_first_index(Xd[cand_t])
    assert np.array_equal(eidx, z[f"{tag}_wk_embed_idx"])
    e_t = torch.as_tensor(eidx, device=dev)
    weight = (1 + w_unc * cw * torch.from_numpy(score).to(dev))[e_t].contiguous()
    kk = min(k, len(eidx))
    assert kk == int(z[f"{tag}_wk_qsize"])
    res = KM.kmeans_fit_select(Xd[cand_t][e_t], kk, sample_weight=weight)
    assert np.array_equal(res.labels.cpu().numpy(), z[f"{tag}_wk_labels"])
    assert [cand[i] for i in res.query_rows] == z[f"{tag}_wk_query"].tolist()
    assert res.n_iter == int(z[f"{tag}_wk_niter"])


@pytest.mark.parametrize("n,d,k,weighted", [(3000, 2048, 150, False), (2500, 2048, 120, True), (700, 96, 333, False),
                                            (1000, 36, 10, True), (130, 2048, 100, False), (5000, 256, 64, False)])
def test_kmeans_matches_sklearn(built_lib, n, d, k, weighted):
    """The estimator itself against sklearn.cluster.KMeans(random_state=318) on fresh pools (odd feature counts,
    k close to n: two-member clusters whose closest member is a rounding-level tie)."""
    from sklearn.cluster import KMeans
    from oracle import vatl_oracle as O
    from vatlq import kmeans as KM
    rng = np.random.default_rng(n + k)
    cen = np.abs(rng.normal(0, 1, (max(n // 25, 1), d)))
    X = (cen[(np.arange(n) // 25) % len(cen)] + rng.normal(0, 0.2, (n, d))).astype(np.float32)
    w = (1 + rng.random(n)) if weighted else None
    X64 = X.astype(np.float64)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        km = KMeans(n_clusters=k, random_state=318)
        lab = km.fit_predict(X64, sample_weight=w)
        rows = O._closest_member_per_cluster(X64, km, lab)
    dev = "cuda:0"
    res = KM.kmeans_fit_select(torch.from_numpy(X).to(dev), k, sample_weight=None if w is None else torch.from_numpy(w).to(dev))
    assert np.array_equal(res.labels.cpu().numpy(), lab)
    assert res.query_rows == rows
    assert res.n_iter == km.n_iter_
    assert np.allclose(res.centers.cpu().numpy(), km.cluster_centers_, rtol=0, atol=1e-9)


@pytest.mark.parametrize("n,nw,k", [(100, 30, 40), (200, 50, 60), (120, 100, 110)])
def test_kmeans_empty_cluster_relocation(built_lib, n, nw, k):
    """More clusters than rows that weigh anything: sklearn re-seeds the empty clusters from the farthest rows
    (_relocate_empty_clusters_dense) and leaves those that stay empty at the biggest cluster's row
    (_average_centers).  The reference's usage never gets here (its weights are >= 1), the estimator does: labels and
    iteration counts against sklearn."""
    from sklearn.cluster import KMeans
    from vatlq import kmeans as KM
    rng = np.random.default_rng(5 + n)
    X = np.abs(rng.normal(0, 1, (n, 16))).astype(np.float32)
    w = np.zeros(n)
    w[rng.choice(n, nw, replace=False)] = 1 + rng.random(nw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        km = KMeans(n_clusters=k, random_state=318)
        lab = km.fit_predict(X.astype(np.float64), sample_weight=w)
    res = KM.kmeans_fit_select(torch.from_numpy(X).cuda(), k, sample_weight=torch.from_numpy(w).cuda(), select=False)
    assert res.relocations > 0
    assert np.array_equal(res.labels.cpu().numpy(), lab) and res.n_iter == km.n_iter_


def test_kmeans_kernels_against_numpy(built_lib):
    """The pieces on their own: numpy-ordered column means and squared distances (bit-equal), seeding distances,
    assignment (first minimum), ragged shapes (n not a multiple of the tile, d % 32 != 0)."""
    import ctypes as C
    v = built_lib
    from vatlq import kmeans as KM
    L = v._lib.lib()
    dev = "cuda:0"
    rng = np.random.default_rng(5)
    n, d, k = 1237, 100, 77
    X = rng.normal(0, 1, (n, d)).astype(np.float32)
    Xd = torch.from_numpy(X).to(dev)
    ws = torch.empty(int(L.vatlq_kmeans_workspace_bytes(n, d, k)), dtype=torch.uint8, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    mean = torch.empty(d, dtype=torch.float64, device=dev)
    mv = torch.empty(1, dtype=torch.float64, device=dev)
    assert L.vatlq_kmeans_mean_var(p(Xd), n, d, p(mean), p(mv), p(ws), ws.numel(), st) == 0
    X64 = X.astype(np.float64)
    assert np.array_equal(mean.cpu().numpy(), X64.mean(axis=0))
    assert np.isclose(mv.item(), np.var(X64, axis=0).mean(), rtol=1e-12)
    Cc = rng.normal(0, 1, (k, d))
    Cc[5] = Cc[3]                                           # duplicate centre: the first one wins
    Cd = torch.from_numpy(Cc).to(dev)
    labels = torch.empty(n, dtype=torch.int32, device=dev)
    old = torch.zeros(n, dtype=torch.int32, device=dev)
    changed = torch.zeros(1, dtype=torch.int32, device=dev)
    assert L.vatlq_kmeans_assign(p(Xd), n, d, p(Cd), k, p(labels), p(old), p(changed), p(ws), ws.numel(), st) == 0
    val = (Cc * Cc).sum(1)[None, :] - 2 * X64 @ Cc.T
    ref = val.argmin(1)
    got = labels.cpu().numpy()
    gap = np.sort(val, axis=1)
    clear = (gap[:, 1] - gap[:, 0]) > 1e-9                  # rows whose winner is not a rounding-level tie
    assert np.array_equal(got[clear], ref[clear]) and (got != 5).all()
    assert int(changed.item()) == int((got != 0).sum())
    dis = torch.empty(n, dtype=torch.float64, device=dev)
    assert L.vatlq_kmeans_rowdist(p(Xd), n, d, p(Cd), p(labels), p(dis), st) == 0
    assert np.array_equal(dis.cpu().numpy(), ((X64 - Cc[got]) ** 2).sum(axis=1))
    # k-means++ seeds against sklearn's own routine on the same stream
    from sklearn.cluster import kmeans_plusplus
    _, idx_ref = kmeans_plusplus(X64 - X64.mean(axis=0), k, random_state=np.random.RandomState(318))
    res = KM.kmeans_fit_select(Xd, k)
    assert res.center_ids.cpu().tolist() == idx_ref.tolist()


def test_kmeans_errors_like_the_reference(built_lib):
    """More clusters than distinct rows: sklearn leaves clusters empty and the reference's per-cluster argmin raises
    ValueError (:602); n_samples < n_clusters raises inside sklearn — same exceptions here."""
    from vatlq import kmeans as KM
    dev = "cuda:0"
    rng = np.random.default_rng(2)
    X = rng.normal(0, 1, (300, 32)).astype(np.float32)
    X[100:200] = X[0:100]
    with pytest.raises(ValueError, match="empty sequence"):
        KM.kmeans_fit_select(torch.from_numpy(X).to(dev), 250)
    with pytest.raises(ValueError, match="should be >= n_clusters"):
        KM.kmeans_fit_select(torch.from_numpy(X[:10]).to(dev), 11)
    with pytest.raises(v_err(built_lib)):
        KM.kmeans_fit_select(torch.from_numpy(X), 3)        # CPU tensor: no CPU path


def v_err(v):
    return v._lib.VatlqError


@pytest.mark.parametrize("flt", ["K-Means", "weighted"])
def test_controller_kmeans_filters(built_lib, flt):
    """ActiveLearning._query with filter K-Means / weighted against the oracle's restatement of the same lines."""
    from oracle import vatl_oracle as O
    v = built_lib
    synth = v.synth
    n = 400
    cfg = SimpleNamespace(VAL=SimpleNamespace(QUERY_RATIO=[0.05, 0.1], W_UNC=0.8, UNC_LAMBDA=0.01),
                          DATA_PRESET=SimpleNamespace(HEATMAP_SIZE=[64, 48]), AE=SimpleNamespace(Z_DIM=4))
    opt = SimpleNamespace(strategy=f"THC+None_{flt}filter", uncertainty="THC", representativeness="None", filter=flt,
                          video_id="0", THCvsWPU="const", fixed_lambda=False, onebyone=False)
    al = v.ActiveLearning(cfg, opt, eval_len=n)
    dev = "cuda:0"
    X = synth.pool_embeddings(n, kind="weak", seed=11)
    X[50:70] = X[10:30]                                     # duplicates for np.unique
    labeled = list(range(0, n, 9))
    al.labeled_id.update(labeled); al.unlabeled_id.difference_update(labeled)
    unl = al.unlabeled_id.index
    score = np.zeros(n)
    score[unl] = synth.pool_unc(n, seed=11)[unl]
    got = al._kmeans_query(torch.from_numpy(X).to(dev), torch.from_numpy(score).to(dev), unl, 0.4)
    cand = sorted(unl)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if flt == "K-Means":
            ref, qs, _ = O.kmeans_filter(X.astype(np.float64), cand, 20, len(unl))
        else:
            ref, qs, _, _ = O.weighted_kmeans_filter(X.astype(np.float64), cand, score[cand], 0.8, 0.4, 20, len(unl))
    assert got == ref and al.query_size == qs
