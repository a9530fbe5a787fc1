"""CPU side of the K-Means / weighted filters (ActiveLearning.py:553-580, 593-608): the oracle against the golden
written from the reference's own statements (oracle/pin_against_reference.py kmeans), and the host pieces of
kmeans.py that need no GPU (the random stream that is handed to the device)."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kmeans_cases import CASES, case_inputs     # noqa: E402


def test_oracle_reproduces_reference_golden():
    from oracle import vatl_oracle as O
    z = np.load(os.path.join(ROOT, "tests", "golden", "kmeans.npz"))
    for tag in CASES:
        X, cand, score, k, w_unc, cw = case_inputs(z, tag)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            q, qs, lab = O.kmeans_filter(X.astype(np.float64), cand, k)
            qw, qsw, labw, eidx = O.weighted_kmeans_filter(X.astype(np.float64), cand, score, w_unc, cw, k)
        assert q == z[f"{tag}_km_query"].tolist() and qs == int(z[f"{tag}_km_qsize"])
        assert np.array_equal(lab, z[f"{tag}_km_labels"])
        assert qw == z[f"{tag}_wk_query"].tolist() and qsw == int(z[f"{tag}_wk_qsize"])
        assert np.array_equal(labw, z[f"{tag}_wk_labels"]) and np.array_equal(eidx, z[f"{tag}_wk_embed_idx"])


def test_random_stream_matches_sklearn_seeding():
    """kmeans.py draws `choice` once and then ONE block of (k-1)*trials uniforms; sklearn draws uniform(size=trials)
    k-1 times.  Same stream — checked by replaying sklearn's _kmeans_plusplus with the block and comparing seeds."""
    from sklearn.cluster import kmeans_plusplus
    rng = np.random.default_rng(0)
    X = rng.normal(0, 1, (300, 16))
    k = 40
    _, idx_ref = kmeans_plusplus(X, k, random_state=np.random.RandomState(318))
    rs = np.random.RandomState(318)
    n = X.shape[0]
    sw = np.ones(n)
    first = int(rs.choice(n, p=sw / sw.sum()))
    trials = 2 + int(np.log(k))
    rand = rs.uniform(size=(k - 1) * trials).reshape(k - 1, trials)
    xx = (X * X).sum(1)
    closest = np.maximum(xx + xx[first] - 2 * X @ X[first], 0)
    pot = closest.sum()
    ids = [first]
    for c in range(1, k):
        cand = np.minimum(np.searchsorted(np.cumsum(closest), rand[c - 1] * pot), n - 1)
        D = np.minimum(np.maximum(xx[cand][:, None] + xx[None, :] - 2 * X[cand] @ X.T, 0), closest[None, :])
        b = int(np.argmin(D.sum(1)))
        pot, closest = D[b].sum(), D[b]
        ids.append(int(cand[b]))
    assert ids == idx_ref.tolist()


def test_device_algorithm_restated_in_numpy_equals_sklearn():
    """oracle/kmeans_restated.py is the device pipeline (kmeans.cu + kmeans.py) step for step in NumPy: k-means++ on the
    host's random block, Lloyd with centred-frame centre sums, relocation of empty clusters, numpy-ordered final
    distances.  It must give sklearn's labels, iteration counts and the reference's queried rows — on the golden
    pools, on pools with many two-member clusters (rounding-level ties), with weights, and with empty clusters."""
    from sklearn.cluster import KMeans
    from oracle import kmeans_restated as KR
    from oracle import vatl_oracle as O
    z = np.load(os.path.join(ROOT, "tests", "golden", "kmeans.npz"))
    for tag in ("clustered", "pairs", "one", "all", "dups"):
        X, cand, score, k, w_unc, cw = case_inputs(z, tag)
        rows, labels, n_iter, _, _ = KR.fit_select(X[cand], k)
        assert [cand[i] for i in rows] == z[f"{tag}_km_query"].tolist()
        assert np.array_equal(labels, z[f"{tag}_km_labels"]) and n_iter == int(z[f"{tag}_km_niter"])
    rng = np.random.default_rng(7)
    for n, d, k, weighted in ((300, 64, 250, False), (400, 32, 60, True), (120, 256, 100, True)):
        X = np.abs(rng.normal(0, 1, (n, d))).astype(np.float32)
        w = (1 + rng.random(n)) if weighted else None
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            km = KMeans(n_clusters=k, random_state=318)
            lab = km.fit_predict(X.astype(np.float64), sample_weight=w)
            ref_rows = O._closest_member_per_cluster(X.astype(np.float64), km, lab)
        rows, labels, n_iter, _, _ = KR.fit_select(X, k, w)
        assert np.array_equal(labels, lab) and n_iter == km.n_iter_ and rows == ref_rows
    # more clusters than rows that weigh anything: relocation + clusters that stay empty
    X = np.abs(rng.normal(0, 1, (100, 16))).astype(np.float32)
    w = np.zeros(100)
    w[rng.choice(100, 30, replace=False)] = 1 + rng.random(30)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        km = KMeans(n_clusters=40, random_state=318)
        lab = km.fit_predict(X.astype(np.float64), sample_weight=w)
    ids = KR.seeds(X.astype(np.float64), w, 40, np.random.RandomState(318))
    labels, _, n_iter, reloc = KR.lloyd(X.astype(np.float64), w, ids, np.var(X.astype(np.float64), axis=0).mean() * 1e-4)
    assert reloc > 0 and np.array_equal(labels, lab) and n_iter == km.n_iter_
