"""Core-set parity at BASELINE scale: the default GPU path (16 picks per round, exact pruning on) must
reproduce the pick lists the REFERENCE's own coreset_selection produced on the same pools
(oracle/pin_scale.py -> tests/golden/coreset_scale_*.npz; the counter-based generators of synth.py give the
same bits in numpy there and in torch here).  Bar: the selected index list is bit-exact, in pick order."""
import hashlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
LAM = 0.01


def _gold(tag):
    f = os.path.join(GOLD, f"coreset_scale_{tag}.npz")
    if not os.path.exists(f):
        pytest.skip(f"{f} not generated yet (oracle/pin_scale.py {tag})")
    return np.load(f)


def _sha(p):
    return hashlib.sha256(np.asarray(p, dtype="<i8").tobytes()).hexdigest()


@pytest.mark.parametrize("tag", ["weak", "iid", "c3", "c3lab"])
def test_coreset_only_matches_reference_at_scale(tag):
    import vatlq
    from vatlq import ops, synth
    z = _gold(tag)
    assert not bool(z["full_query"])
    n, k_full, moks = int(z["n"]), int(z["k_full"]), float(z["moks"])
    dev = torch.device("cuda:0")
    X = synth.pool_embeddings(n, kind=str(z["kind"]), device=dev)
    lab = synth.pool_labeled(n, int(n * float(z["labeled_frac"])))
    unc = synth.pool_unc(n, device=dev)
    if lab.size:
        unc[torch.from_numpy(lab).to(dev)] = 0.0
    ops.set_prune("env")
    ops.prune_stats(reset=True)
    picks, st = ops.coreset_select(X, unc, lab, k_full, moks, LAM, batch=16)     # the default path
    got = picks.cpu().numpy()
    gold = z["picks"]
    assert np.array_equal(got[:len(gold)], gold), f"first difference at pick {int(np.argmax(got[:len(gold)] != gold))}"
    if len(gold) == k_full:
        assert _sha(got) == str(z["sha256"])
    if n >= 8192 and str(z["kind"]) == "clustered":      # the pruned path really ran
        pr = ops.prune_stats(reset=True)
        assert pr["tiles"] > 0 and pr["streamed"] < pr["tiles"]
    assert st.batch == 16


@pytest.mark.parametrize("tag", ["c4", "c4lab", "c5"])
def test_full_query_matches_reference_at_scale(tag):
    """Heat maps -> scan -> WPU -> fusion -> core-set on the bench pools (configs 4 and 5) against the oracle's
    scoring loop + the reference's selection; also pins a few scores of the whole-pool scoring."""
    import vatlq
    from vatlq import synth
    z = _gold(tag)
    assert bool(z["full_query"])
    n, k_full, moks = int(z["n"]), int(z["k_full"]), float(z["moks"])
    free, _ = torch.cuda.mem_get_info()
    need = min(n, synth.HEAT_RING) * synth.FRAME_BYTES + n * 2048 * 4 * 1.2 + (6 << 30)
    if free < need:
        pytest.skip("not enough free HBM for this pool")
    dev = torch.device("cuda:0")
    segs, bb, ip, inx, X, _ = synth.rank_pool(n, 0, n, dev, str(z["kind"]))
    lab = synth.pool_labeled(n, int(n * float(z["labeled_frac"]))).tolist()
    res = vatlq.run_query(segs, bb, ip, inx, X, synth.ae_weights(42, 4), lab, k_full, moks, LAM, batch=16, device=dev)
    got = res.picks.cpu().numpy()
    gold = z["picks"]
    # scoring parity on the pinned head of the pool (tolerance: north_star's 1e-5 relative)
    assert np.allclose(res.thc[:64].cpu().numpy(), z["thc_head"], rtol=1e-5, atol=1e-6)
    assert np.allclose(res.wpu[:64].cpu().numpy(), z["wpu_head"], rtol=1e-5, atol=1e-7)
    assert np.allclose(res.peak_mean[:64].cpu().numpy(), z["peak_head"], rtol=1e-5, equal_nan=True)
    assert np.isclose(float(res.thc.double().sum()), float(z["thc_sum"]), rtol=1e-6)
    assert np.isclose(float(res.wpu.double().sum()), float(z["wpu_sum"]), rtol=1e-5)
    assert np.isclose(res.combine_weight, float(z["combine_weight"]), rtol=1e-6)
    assert int(torch.argmax(res.unc)) == int(z["top_unc_idx"][0])
    assert np.array_equal(got[:len(gold)], gold), f"first difference at pick {int(np.argmax(got[:len(gold)] != gold))}"
    if len(gold) == k_full:
        assert _sha(got) == str(z["sha256"])
    del segs, X, res
    torch.cuda.empty_cache()
