"""GPU parity: k-center greedy core-set.  Selected indices must equal the reference's pick list
bit for bit (ties -> lowest index); distances within 1e-5 relative."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("batch", [1, 8, 16])
def test_golden_coreset(built_lib, gold_coreset, batch):
    v = built_lib
    for tag, c in gold_coreset.items():
        first = -1
        labeled = c["labeled"].tolist()
        picks, st, md, unc_after = v.ops.coreset_select(
            torch.from_numpy(c["X"]).cuda(), torch.from_numpy(c["unc"]).cuda(), labeled, int(c["k"]),
            float(c["moks"]), float(c["lam"]), rule=str(c["rule"]), first_pick=first, batch=batch, return_state=True)
        assert picks.cpu().tolist() == c["picks"].tolist(), (tag, batch)
        # the reference's d(c,c) is ~1e-8 (BLAS dot vs einsum norms), ours is exactly 0
        assert np.allclose(md.cpu().numpy(), c["min_d"], rtol=1e-5, atol=1e-6), tag
        assert (unc_after.cpu().numpy()[c["picks"]] == 0).all()
        assert st.picks == int(c["k"]) and st.passes >= 1
        if batch == 8 and int(c["k"]) >= 40:
            assert st.passes < int(c["k"]), (tag, st)         # batching really saved passes over X


def test_pairwise_distance_parity(built_lib):
    from sklearn.metrics import pairwise_distances
    v = built_lib
    X = v.synth.embeddings(500, d=2048, seed=8)
    cen = [0, 3, 499, 17, 250, 250, 31, 32, 33, 100, 7]
    out = v.ops.pairwise_dist(torch.from_numpy(X).cuda(), cen).cpu().numpy()
    ref = pairwise_distances(X.astype(np.float64), X[cen].astype(np.float64))
    assert np.allclose(out, ref, rtol=1e-5, atol=1e-6)
    far = ref > 1.0
    assert np.abs(out[far] / ref[far] - 1).max() < 1e-12          # fp64 accumulation, not fp32
    assert (out[cen, np.arange(len(cen))] == 0).all()             # d(c,c) is exactly 0 here


def test_reference_style_call(built_lib):
    """coreset_selection(self, embeddings, uncertainty) with the reference's attribute names."""
    from types import SimpleNamespace
    from oracle import vatl_oracle as O
    v = built_lib
    X = v.synth.embeddings(400, d=512, seed=3).astype(np.float64)
    unc = np.random.default_rng(0).uniform(0, 1, 400)
    lab = [5, 77, 300]
    unc[lab] = 0
    fake = SimpleNamespace(labeled_id=v.IndexCollection(lab), moks_queried=0.45, unc_lambda=0.01, uncertainty="THC+WPU",
                           cfg=SimpleNamespace(VAL=SimpleNamespace(UNC_LAMBDA=0.01)), opt=SimpleNamespace(fixed_lambda=False),
                           query_size=35)
    mine_unc = unc.copy()
    picks = v.coreset_selection(fake, X, mine_unc)
    ref_unc = unc.copy()
    ref, _ = O.coreset_select(X, ref_unc, lab, 35, 0.45, 0.01)
    assert picks == ref and np.array_equal(mine_unc, ref_unc)


def test_clustered_precision_hazard(built_lib):
    """clustered embeddings flip picks under fp32 accumulation; fp64 must not."""
    from oracle import vatl_oracle as O
    v = built_lib
    X = v.synth.embeddings(3000, d=2048, seed=2, clustered=True)
    unc = np.random.default_rng(1).uniform(0, 1, 3000)
    ref, _ = O.coreset_select(X.astype(np.float64), unc.copy(), [], 150, 0.6, 0.01)
    for batch in (1, 8):
        picks, st = v.ops.coreset_select(torch.from_numpy(X).cuda(), torch.from_numpy(unc).cuda(), [], 150, 0.6, 0.01, batch=batch)
        assert picks.cpu().tolist() == ref, batch


def test_batched_equals_sequential_at_scale(built_lib):
    """Config 3 scale (100 k x 2048) through a size-independent property: the batched planner
    must return exactly the picks of the one-pick-per-pass form; picks are distinct; min_d is 0
    at the picks."""
    v = built_lib
    n, k = 100000, 160
    X = v.synth.device_embeddings(n, "cuda:0", seed=2)
    unc = torch.rand(n, dtype=torch.float64, device="cuda:0", generator=torch.Generator("cuda:0").manual_seed(5))
    p1, s1 = v.ops.coreset_select(X, unc, [], k, 0.6, 0.01, batch=1)
    p8, s8, md, _ = v.ops.coreset_select(X, unc, [], k, 0.6, 0.01, batch=8, return_state=True)
    assert torch.equal(p1, p8)
    assert len(set(p8.cpu().tolist())) == k
    assert (md[p8] == 0).all() and s8.passes < s1.passes
    lab = p8[:50].cpu().tolist()
    unc2 = unc.clone(); unc2[p8[:50]] = 0
    q1, _ = v.ops.coreset_select(X, unc2, lab, 60, 0.3, 0.01, batch=1)
    q8, _ = v.ops.coreset_select(X, unc2, lab, 60, 0.3, 0.01, batch=8)
    assert torch.equal(q1, q8) and not set(q8.cpu().tolist()) & set(lab)
