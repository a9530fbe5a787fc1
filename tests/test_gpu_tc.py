"""The labelled-set initialisation on the tcgen05 tensor cores (csrc/tc_dist.cu): the TF32 GEMM only
pre-selects (row, centre) pairs under a proved error bound, the listed pairs are re-scored with the
canonical fp64 arithmetic — so min_d must equal the exact passes' min_d BIT FOR BIT, the sampled bound
check must never fire, and the values must agree with sklearn's float64 distances (<= 1e-9 relative)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _exact_init(v, X, lab, lo, hi):
    import ctypes as C
    L = v._lib.lib()
    n, d = X.shape
    ws_bytes = L.vatlq_coreset_workspace_bytes(n, d, 16)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=X.device)
    md = torch.full((n,), -1.0, dtype=torch.float64, device=X.device)
    v._lib.check(L.vatlq_coreset_init(C.c_void_p(X.data_ptr()), n, d, lo, hi, C.c_void_p(lab.data_ptr()), lab.numel(),
                                      C.c_void_p(md.data_ptr()), C.c_void_p(ws.data_ptr()), ws_bytes,
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), "init")
    return md


@pytest.mark.parametrize("kind,n,L,lo,hi", [("clustered", 6000, 600, 0, 6000), ("iid", 4100, 1003, 0, 4100),
                                            ("weak", 5000, 257, 1111, 4321), ("clustered", 20000, 2000, 0, 20000)])
def test_tc_init_equals_exact_passes_bit_for_bit(built_lib, kind, n, L, lo, hi):
    v = built_lib
    dev = torch.device("cuda:0")
    X = v.synth.pool_embeddings(n, kind=kind, device=dev)
    lab = torch.from_numpy(v.synth.pool_labeled(n, L)).to(dev)
    md_tc = torch.full((n,), -1.0, dtype=torch.float64, device=dev)
    ok, st, tmin = v.ops.coreset_init_tc(X, lab, md_tc, lo, hi, verify=True, want_tmin=True)
    assert ok, st
    assert st["violations"] == 0, st                      # |t~ - t| <= E held on every sampled entry
    assert 0 < st["pairs"] <= st["capacity"]
    md_ex = _exact_init(v, X, lab, lo, hi)
    assert torch.equal(md_tc[lo:hi], md_ex[lo:hi]), (md_tc[lo:hi] - md_ex[lo:hi]).abs().max()
    assert bool((md_tc[:lo] == -1).all()) and bool((md_tc[hi:] == -1).all())      # rows outside the range untouched
    # the approximate minimum brackets the exact one within the bound
    t_ex = (md_ex[lo:hi] ** 2).float()
    xx = (X[lo:hi].double() ** 2).sum(1)
    cmax = (X[lab].double() ** 2).sum(1).max().sqrt()
    E = 2.0 / 256.0 * xx.sqrt() * cmax + (xx + cmax ** 2) * 2.0 ** -20
    assert bool(((tmin.double() - t_ex.double()).abs() <= E * 1.001 + 1e-6).all())
    # the list stays short: a handful of centre groups per row
    assert st["pairs"] < 24 * (hi - lo)


def test_tc_init_matches_sklearn(built_lib):
    from sklearn.metrics import pairwise_distances
    v = built_lib
    dev = torch.device("cuda:0")
    n, L = 3000, 300
    Xh = v.synth.pool_embeddings(n, kind="clustered")
    labh = v.synth.pool_labeled(n, L)
    X, lab = torch.from_numpy(Xh).to(dev), torch.from_numpy(labh).to(dev)
    md = torch.empty(n, dtype=torch.float64, device=dev)
    ok, st, _ = v.ops.coreset_init_tc(X, lab, md, 0, n)
    assert ok
    ref = pairwise_distances(Xh.astype(np.float64), Xh[labh].astype(np.float64)).min(axis=1)
    got = md.cpu().numpy()
    far = ref > 1e-3
    assert np.allclose(got[far], ref[far], rtol=1e-9)
    assert np.all(got[labh] < 1e-5)          # a labelled row is at distance ~0 of itself (sklearn: not exactly 0 either)


def test_selection_uses_tc_init_and_picks_are_unchanged(built_lib):
    """ops.coreset_select dispatches the initialisation to the tensor-core path for >= 256 labelled rows;
    the pick list equals the one after the exact passes."""
    v = built_lib
    dev = torch.device("cuda:0")
    n = 12000
    X = v.synth.pool_embeddings(n, kind="clustered", device=dev)
    labh = v.synth.pool_labeled(n, 1200)
    unc = v.synth.pool_unc(n, device=dev)
    unc[torch.from_numpy(labh).to(dev)] = 0
    before = dict(v.ops._tc_init_stats)
    p_tc, _, md_tc, _ = v.ops.coreset_select(X, unc, labh, 300, 0.6, 0.01, return_state=True)
    assert v.ops._tc_init_stats["calls"] == before["calls"] + 1
    p_ex, _, md_ex, _ = v.ops.coreset_select(X, unc, labh, 300, 0.6, 0.01, return_state=True, tc_init=False)
    assert torch.equal(p_tc, p_ex) and torch.equal(md_tc, md_ex)


def test_tc_init_small_and_ragged_shapes(built_lib):
    """Fewer rows than one 128-row tile, a ragged last tile, L not a multiple of 8, duplicate labelled rows."""
    v = built_lib
    dev = torch.device("cuda:0")
    for n, L in ((100, 16), (300, 260), (1000, 257)):
        X = v.synth.pool_embeddings(n, kind="clustered", device=dev)
        lab_h = v.synth.pool_labeled(n, L)
        if L >= 2:
            lab_h[1] = lab_h[0]            # a duplicate centre
        lab = torch.from_numpy(lab_h).to(dev)
        md = torch.full((n,), -1.0, dtype=torch.float64, device=dev)
        ok, st, _ = v.ops.coreset_init_tc(X, lab, md, 0, n, verify=True)
        assert ok and st["violations"] == 0
        assert torch.equal(md, _exact_init(v, X, lab, 0, n))
