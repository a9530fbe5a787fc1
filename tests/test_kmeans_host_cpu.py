"""CPU: the host side of the K-Means filter (vatl4pose-wacv2024_b200/kmeans.py — sklearn's random stream, the Lloyd
control flow, empty-cluster relocation, the per-cluster pick and its error behaviour) executed end to end with NumPy
stand-ins for the ten vatlq_kmeans_* device calls.  The stand-ins read and write the caller's tensors through the raw
pointers kmeans.py hands to the C ABI, so the argument order / buffer sizes of every call are exercised too.  The
device kernels themselves are checked on the GPU (tests/test_gpu_kmeans.py); the product never runs these stand-ins."""
import contextlib
import ctypes as C
import os
import sys
import warnings

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kmeans_cases import case_inputs     # noqa: E402


def _a(ptr, count, ctype):
    """NumPy view of `count` items at a c_void_p the code under test passed to the 'library'."""
    addr = ptr.value if isinstance(ptr, C.c_void_p) else int(ptr)
    return np.ctypeslib.as_array(C.cast(addr, C.POINTER(ctype)), shape=(int(count),))


class FakeLib:
    """NumPy restatements of the device calls (same semantics as csrc/kmeans.cu)."""

    def vatlq_kmeans_workspace_bytes(self, n, d, k):
        return 64

    def vatlq_kmeans_mean_var(self, X, n, d, mean, out1, ws, wsb, st):
        x = _a(X, n * d, C.c_float).reshape(n, d).astype(np.float64)
        _a(mean, d, C.c_double)[:] = x.mean(axis=0)
        _a(out1, 1, C.c_double)[0] = np.var(x, axis=0).mean()
        return 0

    def vatlq_kmeans_pp(self, X, n, d, w, k, first, rand, trials, center_ids, closest, ws, wsb, st):
        x = _a(X, n * d, C.c_float).reshape(n, d).astype(np.float64)
        wv = np.ones(n) if w is None else _a(w, n, C.c_double)
        rv = _a(rand, (k - 1) * trials, C.c_double).reshape(k - 1, trials) if k > 1 else None
        xx = (x * x).sum(1)
        ids = [int(first)]
        cl = np.maximum((-2 * (x @ x[first]) + xx[first]) + xx, 0)
        pot = (cl * wv).sum()
        for c in range(1, k):
            cand = np.minimum(np.searchsorted(np.cumsum(wv * cl), rv[c - 1] * pot), n - 1)
            D = np.minimum(np.maximum((-2 * (x[cand] @ x.T) + xx[cand][:, None]) + xx[None, :], 0), cl[None, :])
            pots = D @ wv
            b = int(np.argmin(pots))
            pot, cl = pots[b], D[b]
            ids.append(int(cand[b]))
        _a(center_ids, k, C.c_int32)[:] = ids
        _a(closest, n, C.c_double)[:] = cl
        return 0

    def vatlq_kmeans_gather(self, X, d, ids, k, mean, Cc, Cr, st):
        i = _a(ids, k, C.c_int32)
        n = int(i.max()) + 1
        x = _a(X, n * d, C.c_float).reshape(n, d).astype(np.float64)
        m = _a(mean, d, C.c_double)
        cc = x[i] - m
        _a(Cc, k * d, C.c_double)[:] = cc.reshape(-1)
        _a(Cr, k * d, C.c_double)[:] = (cc + m).reshape(-1)
        return 0

    def vatlq_kmeans_assign(self, X, n, d, Cr, k, labels, labels_old, changed, ws, wsb, st):
        x = _a(X, n * d, C.c_float).reshape(n, d).astype(np.float64)
        c = _a(Cr, k * d, C.c_double).reshape(k, d)
        lab = ((c * c).sum(1)[None, :] - 2 * x @ c.T).argmin(1).astype(np.int32)
        if changed is not None:
            old = _a(labels_old, n, C.c_int32)
            _a(changed, 1, C.c_int32)[0] = int((lab != old).sum())
        _a(labels, n, C.c_int32)[:] = lab
        return 0

    def vatlq_kmeans_update(self, X, n, d, w, mean, labels, k, sums, wsum, order, starts, n_empty, ws, wsb, st):
        x = _a(X, n * d, C.c_float).reshape(n, d).astype(np.float64)
        wv = np.ones(n) if w is None else _a(w, n, C.c_double)
        m = _a(mean, d, C.c_double)
        lab = _a(labels, n, C.c_int32)
        s, ws_ = np.zeros((k, d)), np.zeros(k)
        np.add.at(s, lab, (x - m) * wv[:, None])
        np.add.at(ws_, lab, wv)
        o = np.argsort(lab, kind="stable").astype(np.int32)
        _a(order, n, C.c_int32)[:] = o
        _a(starts, k + 1, C.c_int32)[:] = np.searchsorted(lab[o], np.arange(k + 1))
        _a(sums, k * d, C.c_double)[:] = s.reshape(-1)
        _a(wsum, k, C.c_double)[:] = ws_
        _a(n_empty, 1, C.c_int32)[0] = int((ws_ == 0).sum())
        return 0

    def vatlq_kmeans_relocate(self, X, d, w, mean, labels, empty_ids, far_ids, n_empty, sums, wsum, st):
        far = _a(far_ids, n_empty, C.c_int32)
        emp = _a(empty_ids, n_empty, C.c_int32)
        n = self.n
        x = _a(X, n * d, C.c_float).reshape(n, d).astype(np.float64)
        wv = np.ones(n) if w is None else _a(w, n, C.c_double)
        m = _a(mean, d, C.c_double)
        lab = _a(labels, n, C.c_int32)
        k = self.k
        s = _a(sums, k * d, C.c_double).reshape(k, d)
        ws_ = _a(wsum, k, C.c_double)
        for e, f in zip(emp, far):
            o = lab[f]
            s[o] -= (x[f] - m) * wv[f]
            s[e] = (x[f] - m) * wv[f]
            ws_[e] = wv[f]
            ws_[o] -= wv[f]
        return 0

    def vatlq_kmeans_average(self, sums, wsum, k, d, argmax_w, mean, Cc_old, Cc_new, Cr_new, shift, st):
        s = _a(sums, k * d, C.c_double).reshape(k, d)
        ws_ = _a(wsum, k, C.c_double)
        m = _a(mean, d, C.c_double)
        old = _a(Cc_old, k * d, C.c_double).reshape(k, d)
        new = np.empty((k, d))
        for j in range(k):
            if ws_[j] > 0:
                new[j] = s[j] * (1.0 / ws_[j])
            elif argmax_w >= 0:
                new[j] = s[argmax_w] * (1.0 / ws_[argmax_w]) if argmax_w < j else s[argmax_w]
            else:
                new[j] = s[j]
        _a(Cc_new, k * d, C.c_double)[:] = new.reshape(-1)
        _a(Cr_new, k * d, C.c_double)[:] = (new + m).reshape(-1)
        _a(shift, k, C.c_double)[:] = np.sqrt(((new - old) ** 2).sum(1))
        return 0

    def vatlq_kmeans_rowdist(self, X, n, d, Cr, labels, dis, st):
        x = _a(X, n * d, C.c_float).reshape(n, d).astype(np.float64)
        lab = _a(labels, n, C.c_int32)
        c = _a(Cr, self.k * d, C.c_double).reshape(self.k, d)
        _a(dis, n, C.c_double)[:] = ((x - c[lab]) ** 2).sum(axis=1)
        return 0

    def vatlq_kmeans_pick(self, dis, order, starts, k, picks, st):
        n = self.n
        dv, o, s = _a(dis, n, C.c_double), _a(order, n, C.c_int32), _a(starts, k + 1, C.c_int32)
        out = _a(picks, k, C.c_int32)
        for j in range(k):
            mem = o[s[j]:s[j + 1]]
            out[j] = -1 if mem.size == 0 else int(mem[np.argmin(dv[mem])])
        return 0


@pytest.fixture
def km_on_cpu(monkeypatch):
    import vatlq                                        # noqa: F401  (package alias)
    from vatlq import _lib, kmeans as KM
    fake = FakeLib()
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    monkeypatch.setattr(KM, "_cuda", lambda t, dt, name: t.contiguous())
    monkeypatch.setattr(KM, "_stream", lambda: None)
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())

    def run(X, k, w=None, **kw):
        fake.n, fake.k = X.shape[0], k
        return KM.kmeans_fit_select(torch.from_numpy(np.ascontiguousarray(X)), k,
                                    sample_weight=None if w is None else torch.from_numpy(w), **kw)
    return run


def test_host_flow_reproduces_reference_golden(km_on_cpu):
    z = np.load(os.path.join(ROOT, "tests", "golden", "kmeans.npz"))
    for tag in ("clustered", "pairs", "one", "all", "dups"):
        X, cand, score, k, w_unc, cw = case_inputs(z, tag)
        res = km_on_cpu(X[cand], k)
        assert [cand[i] for i in res.query_rows] == z[f"{tag}_km_query"].tolist()
        assert np.array_equal(res.labels.numpy(), z[f"{tag}_km_labels"]) and res.n_iter == int(z[f"{tag}_km_niter"])


def test_host_flow_weights_relocation_and_errors(km_on_cpu):
    from sklearn.cluster import KMeans
    rng = np.random.default_rng(11)
    X = np.abs(rng.normal(0, 1, (300, 24))).astype(np.float32)
    w = 1 + rng.random(300)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        km = KMeans(n_clusters=40, random_state=318)
        lab = km.fit_predict(X.astype(np.float64), sample_weight=w)
    res = km_on_cpu(X, 40, w)
    assert np.array_equal(res.labels.numpy(), lab) and res.n_iter == km.n_iter_ and res.relocations == 0
    # more clusters than rows that weigh anything: relocation, clusters that stay empty, select=False
    X = np.abs(rng.normal(0, 1, (100, 16))).astype(np.float32)
    w = np.zeros(100)
    w[rng.choice(100, 30, replace=False)] = 1 + rng.random(30)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        km = KMeans(n_clusters=40, random_state=318)
        lab = km.fit_predict(X.astype(np.float64), sample_weight=w)
    res = km_on_cpu(X, 40, w, select=False)
    assert res.relocations > 0 and np.array_equal(res.labels.numpy(), lab) and res.n_iter == km.n_iter_
    from oracle import vatl_oracle as O
    assert km_on_cpu(X, 40, w).query_rows == O._closest_member_per_cluster(X.astype(np.float64), km, lab)   # labels 0..29 all used
    # more clusters than distinct rows: some label below cluster_num has no member -> the reference's per-cluster
    # argmin raises (:602), and so does the host flow
    Xd = np.abs(rng.normal(0, 1, (300, 32))).astype(np.float32)
    Xd[100:200] = Xd[0:100]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        km = KMeans(n_clusters=250, random_state=318)
        lab = km.fit_predict(Xd.astype(np.float64))
        with pytest.raises(ValueError):
            O._closest_member_per_cluster(Xd.astype(np.float64), km, lab)
    with pytest.raises(ValueError, match="empty sequence"):
        km_on_cpu(Xd, 250)
    with pytest.raises(ValueError, match="should be >= n_clusters"):
        km_on_cpu(X[:5], 6)


def test_unique_rows_first_index_equals_numpy_unique(monkeypatch):
    """np.unique(E, axis=0, return_index=True)[1] (ActiveLearning.py:555) from a sort by the first column plus a full
    comparison of the runs that share it: duplicates, rows that differ only in later columns, signed zeros."""
    import vatlq  # noqa: F401
    from vatlq import kmeans as KM
    monkeypatch.setattr(KM, "_cuda", lambda t, dt, name: t.contiguous())

    def rank_scores(score, mask=None, descending=True, count=None):        # vatlq_rank_scores: stable, ties by id
        s = np.where(score.numpy() == 0, 0.0, score.numpy())
        order = np.argsort(-s if descending else s, kind="stable")
        return torch.from_numpy(order if count is None else order[:count])
    monkeypatch.setattr(KM, "rank_scores", rank_scores)
    rng = np.random.default_rng(2)
    E = np.abs(rng.normal(0, 1, (200, 12))).astype(np.float32)
    E[50:80] = E[10:40]                           # duplicates
    E[100:110, 0] = E[5, 0]                       # same first column, different rows
    E[120:125, :3] = E[6, :3]                     # same first three columns
    E[130, 0] = 0.0
    E[131, 0] = -0.0                              # signed zeros compare equal
    E[131, 1:] = E[130, 1:]
    got = KM.unique_rows_first_index(torch.from_numpy(E))
    _, ref = np.unique(E.astype(np.float64), axis=0, return_index=True)
    assert np.array_equal(got, ref)
    assert KM.unique_rows_first_index(torch.zeros((0, 4))).size == 0


@pytest.mark.parametrize("flt", ["K-Means", "weighted"])
def test_controller_kmeans_filters_host_logic(km_on_cpu, monkeypatch, flt):
    """ActiveLearning._kmeans_query (candidate list, np.unique de-duplication, weights 1 + w_unc * cw * score, query-size
    clamp, the reference's index mapping of :580) against the oracle's restatement of ActiveLearning.py:553-608."""
    from types import SimpleNamespace
    import vatlq
    from vatlq import _lib, kmeans as KM
    from oracle import vatl_oracle as O
    synth = vatlq.synth

    def rank_scores(score, mask=None, descending=True, count=None):
        s = np.where(score.numpy() == 0, 0.0, score.numpy())
        order = np.argsort(-s if descending else s, kind="stable")
        return torch.from_numpy(order if count is None else order[:count])
    monkeypatch.setattr(KM, "rank_scores", rank_scores)
    n = 300
    cfg = SimpleNamespace(VAL=SimpleNamespace(QUERY_RATIO=[0.05, 0.1], W_UNC=0.8, UNC_LAMBDA=0.01),
                          DATA_PRESET=SimpleNamespace(HEATMAP_SIZE=[64, 48]), AE=SimpleNamespace(Z_DIM=4))
    opt = SimpleNamespace(strategy=f"THC+None_{flt}filter", uncertainty="THC", representativeness="None", filter=flt,
                          video_id="0", THCvsWPU="const", fixed_lambda=False, onebyone=False)
    al = vatlq.ActiveLearning(cfg, opt, eval_len=n)
    X = synth.pool_embeddings(n, kind="weak", seed=11)
    X[50:70] = X[10:30]
    labeled = list(range(0, n, 9))
    al.labeled_id.update(labeled)
    al.unlabeled_id.difference_update(labeled)
    unl = al.unlabeled_id.index
    score = np.zeros(n)
    score[unl] = synth.pool_unc(n, seed=11)[unl]
    fake = _lib.lib()
    cand = sorted(unl)
    m = len(cand) if flt == "K-Means" else len(np.unique(X[cand].astype(np.float64), axis=0))
    fake.n, fake.k = m, min(al.query_size, m)
    got = al._kmeans_query(torch.from_numpy(X), torch.from_numpy(score), unl, 0.4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if flt == "K-Means":
            ref, qs, _ = O.kmeans_filter(X.astype(np.float64), cand, 15, len(unl))
        else:
            ref, qs, _, _ = O.weighted_kmeans_filter(X.astype(np.float64), cand, score[cand], 0.8, 0.4, 15, len(unl))
    assert got == ref and al.query_size == qs
