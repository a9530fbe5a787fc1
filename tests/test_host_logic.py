"""CPU: the C-ABI library loads and exports what include/vatlq.h declares; host-side logic;
the product path refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vatlq.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vatlq_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_lib):
    from vatlq import _lib
    names = declared_symbols()
    assert len(names) >= 25
    h = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/vatlq.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)
    assert _lib.lib().vatlq_abi_version() == 1
    assert _lib.lib().vatlq_wpu_weight_count(42, 4) == 2948
    assert _lib.lib().vatlq_wpu_weight_count(38, 4) == 2752
    assert _lib.lib().vatlq_heatmap_scan_workspace_bytes(256, 17) > 256 * 17 * 24


def test_argument_errors_are_reported_not_hidden(built_lib):
    from vatlq import _lib
    L = _lib.lib()
    rc = L.vatlq_wpu(None, None, None, 38, 4, 0, None, None, None, 10, None)
    assert rc == -1 and b"in_dim" in L.vatlq_last_error()
    rc = L.vatlq_coreset_select(None, 10, 6, 0, 10, None, None, 0, 0.0, 0.0, 0, -1, 1, 8, None, None, None, 0, None, None)
    assert rc == -1
    with pytest.raises(_lib.VatlqError):
        _lib.check(rc, "vatlq_coreset_select")


def test_no_cpu_fallback(built_lib):
    import vatlq
    H = torch.zeros((2, 17, 64, 48))
    with pytest.raises(vatlq._lib.VatlqError):
        vatlq.ops.heatmap_scan(H)
    with pytest.raises(vatlq._lib.VatlqError):
        vatlq.ops.coreset_select(torch.zeros((4, 8)), torch.zeros(4, dtype=torch.float64), [], 1, 0.0, 0.01)
    if not torch.cuda.is_available():
        with pytest.raises(vatlq._lib.VatlqError):
            vatlq.localpeak_mean(np.zeros((17, 64, 48), np.float32))
        with pytest.raises(vatlq._lib.VatlqError):
            vatlq.QueryPass(4, "cpu")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vatl4pose-wacv2024_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no CPU", ""), f"{fn} mentions the oracle"


def test_index_collection_semantics():
    # ALiPy/test/test_indexcollection.py:21-37 style
    from vatlq import IndexCollection
    a = IndexCollection([1, 2, 3])
    a.add(4); a.add(2)
    assert a.index == [1, 2, 3, 4] and 3 in a and 9 not in a
    a.discard(2)
    assert a.index == [1, 3, 4]
    a.update([7, 1, 8])
    assert a.index == [1, 3, 4, 7, 8]
    a.difference_update([3, 8, 100])
    assert a.index == [1, 4, 7] and len(a) == 3
    i = a.index; i.append(5)
    assert a.index == [1, 4, 7]          # .index hands out a copy


def _cfg_opt(unc="THC+WPU", rep="None", flt="Coreset"):
    cfg = SimpleNamespace(VAL=SimpleNamespace(QUERY_RATIO=[0.05, 0.1, 0.2], W_UNC=1.0, UNC_LAMBDA=0.01),
                          DATA_PRESET=SimpleNamespace(HEATMAP_SIZE=[64, 48]), AE=SimpleNamespace(Z_DIM=4))
    opt = SimpleNamespace(strategy=f"{unc}+{rep}_{flt}filter", uncertainty=unc, representativeness=rep, filter=flt,
                          video_id="000000", THCvsWPU="const", fixed_lambda=False, onebyone=False)
    return cfg, opt


def test_strategy_names_and_errors():
    from vatlq import ActiveLearning
    cfg, opt = _cfg_opt()
    al = ActiveLearning(cfg, opt, eval_len=200)
    assert al.query_sizes == [10, 20, 40] and al.query_size == 10 and len(al.unlabeled_id) == 200
    for bad in (("Bogus", "None", "Coreset"), ("THC", "Nope", "Coreset"), ("THC", "None", "Nope")):
        with pytest.raises(ValueError):
            ActiveLearning(*_cfg_opt(*bad), eval_len=10)
    # names that exist in the reference but are not accelerated dispatch back to the reference
    for u in ("VL4Pose",):
        al = ActiveLearning(*_cfg_opt(u), eval_len=10)
        with pytest.raises(NotImplementedError):
            al._require_accelerated()
    for f in ("weighted", "K-Means"):       # device K-Means filters (kmeans.py / csrc/kmeans.cu)
        ActiveLearning(*_cfg_opt("THC", "None", f), eval_len=10)._require_accelerated()
    for u in ("THC", "THC_L1", "WPU", "WPU_hybrid", "THC+WPU", "None", "HP", "TPC", "Entropy", "MPE", "Margin"):
        ActiveLearning(*_cfg_opt(u), eval_len=10)._require_accelerated()
    for r, f in (("Influence", "None"), ("Random", "Diversity"), ("None", "Random"), ("Influence", "Coreset")):
        ActiveLearning(*_cfg_opt("THC", r, f), eval_len=10)._require_accelerated()


def test_outcome_bookkeeping():
    from vatlq import ActiveLearning
    cfg, opt = _cfg_opt()
    al = ActiveLearning(cfg, opt, eval_len=200)
    al.labeled_id.update(range(10)); al.unlabeled_id.difference_update(range(10))
    assert al.outcome() is None and al.round_cnt == 1 and al.query_size == 10      # 20 - 10
    al.labeled_id.update(range(10, 20)); al.unlabeled_id.difference_update(range(10, 20))
    assert al.outcome() is None and al.query_size == 20                            # 40 - 20
    al.labeled_id.update(range(20, 40)); al.unlabeled_id.difference_update(range(20, 40))
    assert al.outcome() is None and al.query_size == 160                           # last round: everything
    al.is_early_stop = True
    al.percentage, al.uncertainty_mean, al.combine_weight, al.moksQ_list = [0.0], [1.5], [0.4], [0.7]
    al.performance, al.performance_ann = [None], [None]
    al.ospa_list, al.ospa_list_ann = [None], [None]
    out = al.outcome()
    # the reference's 20-tuple (ActiveLearning.py:203), per-round lists padded on early stop (:169-178)
    assert isinstance(out, tuple) and len(out) == 20
    assert out[0] is al.percentage and out[3] is al.query_list_list and out[19] is al.moksQ_list
    assert len(al.performance) == len(cfg.VAL.QUERY_RATIO) + 1
    assert len(al.uncertainty_mean) == len(al.moksQ_list) == len(al.combine_weight) == len(al.performance)
    assert al.moksQ_list[-1] == 0.7 and out[14] == 100


def test_oks_bookkeeping_sets_moks_and_early_stop():
    """get_retrain_id / is_finished / get_corresponding_id on given per-item OKS values
    (ActiveLearning.py:707-725, 852-884)."""
    from vatlq import ActiveLearning
    cfg, opt = _cfg_opt()
    opt.retrain_thresh = 0.8
    al = ActiveLearning(cfg, opt, eval_len=10)
    al.labeled_id.update([0, 1]); al.unlabeled_id.difference_update([0, 1])
    oks = {i: v for i, v in enumerate([0.95, 0.5, 0.9, 0.7, 0.85, 0.99, 0.6, 0.9, 0.9, 0.9])}
    retrain, moks = al.get_retrain_id([3, 4], oks)
    assert retrain == [1, 3, 4] and abs(moks - np.mean([0.85, 0.7])) < 1e-15
    assert al.get_corresponding_id(oks, true=True, labeled=True) == [0]
    assert al.get_corresponding_id(oks, true=False, labeled=False) == [3, 4, 6]   # 0.85 < 0.8 + 0.05 in binary
    assert al.is_finished([3, 4], oks) == (100, 100, 100)
    hi = {i: 0.9 for i in range(10)}
    assert al.is_finished([3, 4], hi) == (20.0, 20.0, 20.0)


def test_shard_ranges():
    from vatlq.dist import shard_range, shard_sizes
    for n in (0, 1, 7, 8, 170000, 1000003):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            s = shard_sizes(n, w)
            assert max(s) - min(s) <= 1 and sum(s) == n


def test_ae_weight_packing_layout():
    from vatlq import WholeBodyAE, ops, synth
    W = synth.ae_weights(42, 4)
    flat, ind, z = ops.pack_ae_weights(W, "cpu")
    assert (ind, z) == (42, 4) and flat.numel() == 2948
    assert torch.equal(flat[:42], torch.from_numpy(W[0][0][0]))          # first row of W1
    assert torch.equal(flat[24 * 42:24 * 42 + 24], torch.from_numpy(W[0][1]))   # then b1
    ae = WholeBodyAE(z_dim=4)
    flat2, ind2, z2 = ops.pack_ae_weights(ae.state_dict(), "cpu")
    assert flat2.numel() == 2948 and torch.equal(flat2[:42], ae.encoder[0].weight[0].detach())


def test_synthetic_generators_are_seeded():
    from vatlq import synth
    rng = np.random.default_rng(0)
    ids, ip, inx = synth.track_flags(64, rng, 5.0)
    assert ip[0] == 0 and inx[-1] == 0 and np.array_equal(ip[1:], inx[:-1])
    a, b = synth.heatmaps(3, seed=5), synth.heatmaps(3, seed=5)
    assert a.shape == (3, 17, 64, 48) and a.dtype == np.float32 and np.array_equal(a, b)
    assert (a < 0).any()
    assert synth.embeddings(40, 64).shape == (40, 64)


def test_forward_with_embedding_runs_the_backbone_once():
    """Producer side of SURVEY 8f-4: estimators shaped like the reference's (forward and get_embedding both start with
    self.preact, fastpose.py:44-73) give heat maps AND embedding from one backbone pass; the result equals the
    reference's two calls; other estimators keep the two calls."""
    import torch
    from vatlq import active_learning as AL

    class Pose(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.preact = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.ReLU())
            self.avgpool = torch.nn.AdaptiveAvgPool2d(1)
            self.head = torch.nn.Conv2d(8, 17, 1)
            self.calls = 0
            self.preact.register_forward_hook(lambda *a: setattr(self, "calls", self.calls + 1))

        def forward(self, x):
            return self.head(self.preact(x))

        def get_embedding(self, x):
            return torch.flatten(self.avgpool(self.preact(x)), 1)

    torch.manual_seed(0)
    m = Pose().eval()
    x = torch.randn(5, 3, 16, 12)
    with torch.no_grad():
        out, emb = AL.forward_with_embedding(m, x, True)
        assert m.calls == 1 and len(m.preact._forward_hooks) == 1          # one backbone pass; the temporary hook is gone
        assert torch.equal(out, m(x)) and torch.equal(emb, m.get_embedding(x))
        m.calls = 0
        out2, emb2 = AL.forward_with_embedding(m, x, True, fuse=False)
        assert m.calls == 2 and torch.equal(out2, out) and torch.equal(emb2, emb)
        m.calls = 0
        out3, none = AL.forward_with_embedding(m, x, False)
        assert m.calls == 1 and none is None and torch.equal(out3, out)
        class Wrap(torch.nn.Module):            # `.module` unwrapping (torch.nn.DataParallel on one device)
            def __init__(self, inner):
                super().__init__()
                self.module = inner
            def forward(self, x):
                return self.module(x)
        m.calls = 0
        out4, emb4 = AL.forward_with_embedding(Wrap(m), x, True)
        assert m.calls == 1 and torch.equal(out4, out) and torch.equal(emb4, emb)

    class Plain(torch.nn.Module):           # no preact: the reference's two calls
        def forward(self, x):
            return x[:, :1] * 2
        def get_embedding(self, x):
            return x.flatten(1)[:, :4]
    p = Plain()
    o, e = AL.forward_with_embedding(p, x, True)
    assert torch.equal(o, x[:, :1] * 2) and torch.equal(e, x.flatten(1)[:, :4])
