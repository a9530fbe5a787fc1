"""GPU: the whole query through the reference-shaped controller, against the oracle pipeline."""
import warnings
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class FakeEstimator(torch.nn.Module):
    """Stands in for the pose estimator (out of scope): returns pre-computed heat maps / features
    for the item ids smuggled in the first pixel of every crop."""

    def __init__(self, H, X):
        super().__init__()
        self.H, self.X = H, X

    def forward(self, crops):
        return self.H[crops[:, 0, 0, 0].long()]

    def get_embedding(self, crops):
        return self.X[crops[:, 0, 0, 0].long()]


def loader(n, boxes, ip, inx, bs, gt=None):
    for a in range(0, n, bs):
        b = min(n, a + bs)
        idxs = list(range(a, b))
        inps = torch.zeros((b - a, 3, 1, 2, 2))
        inps[:, :, 0, 0, 0] = torch.arange(a, b, dtype=torch.float32)[:, None]
        inps[:, 1, 0, 0, 0] -= 1
        inps[:, 2, 0, 0, 0] += 1
        inps.clamp_(0, n - 1)
        yield (idxs, inps, None, None, None if gt is None else gt[a:b], None, None, boxes[a:b], boxes[a:b], ip[a:b], inx[a:b])


def cfg_opt(unc, flt):
    cfg = SimpleNamespace(VAL=SimpleNamespace(QUERY_RATIO=[0.05, 0.1, 0.2], W_UNC=1.0, UNC_LAMBDA=0.01),
                          DATA_PRESET=SimpleNamespace(HEATMAP_SIZE=[64, 48]), AE=SimpleNamespace(Z_DIM=4))
    opt = SimpleNamespace(strategy=f"{unc}+None_{flt}filter", uncertainty=unc, representativeness="None", filter=flt,
                          video_id="0", THCvsWPU="const", fixed_lambda=False, onebyone=False)
    return cfg, opt


@pytest.mark.parametrize("unc,flt", [("THC+WPU", "Coreset"), ("THC", "None"), ("WPU", "Coreset")])
def test_two_rounds_match_oracle(built_lib, unc, flt):
    from oracle import vatl_oracle as O
    v = built_lib
    n = 240
    ids, ip, inx = v.synth.track_flags(n, np.random.default_rng(1), 10.0)
    H = v.synth.heatmaps(n, seed=1, track_ids=ids)
    boxes = v.synth.boxes_xyxy(n, 1)
    X = v.synth.embeddings(n, d=2048, seed=2)
    W = v.synth.ae_weights(42, 4)
    cfg, opt = cfg_opt(unc, flt)
    est = FakeEstimator(torch.from_numpy(H).cuda(), torch.from_numpy(X).cuda())
    al = v.ActiveLearning(cfg, opt, model=est, eval_loader=None, eval_len=n, AE=W)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = O.score_pool(H, boxes, ip, inx, O.make_autoencoder(W), drop_ears=(unc == "WPU"))
    # ground truth = the reference's predicted pose + noise, a few joints invisible, one item with none visible:
    # the controller derives OKS (al_metric.py:42-69) and from it moks_queried (ActiveLearning.py:852-858)
    rng = np.random.default_rng(7)
    gt = ref["kpts"].reshape(n, 17, 3).astype(np.float32).copy()
    gt[:, :, :2] += rng.normal(0, 6.0, (n, 17, 2)).astype(np.float32)
    gt[:, :, 2] = (rng.random((n, 17)) < 0.8).astype(np.float32) * 2
    gt[5, :, 2] = 0
    oks_ref = {i: float(O.compute_oks(O.xyxy_to_xywh(boxes[i].tolist()), ref["kpts"][i].astype(np.float32).astype(np.float64),
                                      gt[i].reshape(-1).astype(np.float64))) for i in range(n)}
    labeled, moks = [], 0.0
    for rnd in range(2):
        al.eval_loader = loader(n, boxes, ip, inx, 64, gt)
        assert np.isclose(al.moks_queried, moks, rtol=1e-6)
        al.eval_and_query()
        assert np.allclose([al.OKS_dict[i] for i in range(n)], [oks_ref[i] for i in range(n)], rtol=1e-5, atol=1e-9)
        unl = [i for i in range(n) if i not in labeled]
        if unc == "THC+WPU":
            score = O.fuse_scores(ref["thc"][unl], ref["wpu"][unl], "const")
        else:
            score = O.fuse_scores((ref["thc"] if unc == "THC" else ref["wpu"])[unl])
        k = al.query_sizes[rnd] - len(labeled)
        if flt == "Coreset":
            u = np.zeros(n); u[unl] = score
            expect, _ = O.coreset_select(X.astype(np.float64), u, labeled, k, moks, 0.01)
        else:
            order = sorted(range(len(unl)), key=lambda t: score[t], reverse=True)
            expect = sorted(unl[t] for t in order[:k])
        got = al.query_list_list[f"Round{rnd}"]
        assert got == expect, (rnd, got, expect)
        cw_ref = np.nanmean(ref["peak"][unl]) if unl else None
        assert np.isclose(al.combine_weight[-1], cw_ref, rtol=1e-5)
        if unc == "THC+WPU":
            d = al.uncertainty_dict[f"Round{rnd}"]
            assert np.allclose([d[i][0] for i in range(n)], ref["thc"], rtol=1e-5)
            assert np.allclose([d[i][1] for i in range(n)], ref["wpu"], rtol=1e-5)
        labeled = labeled + expect
        assert al.labeled_id.index == labeled
        moks = float(O.mean_oks_of_queries(expect, oks_ref))
        assert np.isclose(al.moksQ_list[-1], moks, rtol=1e-6) and len(al.moksQ_list) == rnd + 1
        assert al.outcome() is None


def test_run_query_host_inputs(built_lib):
    """The functional API bench.py times end to end: host arrays in, picks out."""
    import __graft_entry__ as g
    g.smoke()
