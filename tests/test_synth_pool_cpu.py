"""CPU: the counter-based pool generators (synth.pool_*) are pure functions of (seed, global index): numpy and torch
produce the same bits, a rank's slice equals the slice of the global pool, the heat-map content has the documented
period, and the scale goldens (oracle/pin_scale.py) carry consistent metadata."""
import glob
import hashlib
import os

import numpy as np
import torch

from conftest import GOLD


def test_numpy_and_torch_generate_identical_bits():
    from vatlq import synth
    n = 5000
    for kind in ("clustered", "weak", "iid"):
        a = synth.pool_embeddings(n, 100, 400, d=256, kind=kind)
        b = synth.pool_embeddings(n, 100, 400, d=256, kind=kind, device="cpu").numpy()
        assert a.dtype == np.float32 and np.array_equal(a, b)
        assert np.array_equal(a, synth.pool_embeddings(n, 0, 400, d=256, kind=kind)[100:])      # slice == global
    assert np.array_equal(synth.pool_unc(n, 7, 99), synth.pool_unc(n, 7, 99, device="cpu").numpy())
    assert np.array_equal(synth.pool_boxes(n, 7, 99), synth.pool_boxes(n, 7, 99, device="cpu").numpy())
    tid, pos, ip, inx = synth.pool_tracks(600, seed=0, ring=256)
    h = synth.pool_heatmaps(tid, pos, 250, 262, ring=256)
    ht = synth.pool_heatmaps(tid, pos, 250, 262, ring=256, device="cpu").numpy()
    assert h.shape == (12, 17, 64, 48) and h.dtype == np.float32 and np.array_equal(h, ht)
    # content period = ring, tracks broken at ring multiples, values exact multiples of 2^-20 in the features
    assert np.array_equal(h[6], synth.pool_heatmaps(tid, pos, 0, 1, ring=256)[0]) and ip[256] == 0 and inx[255] == 0
    x = synth.pool_embeddings(n, 0, 64, d=64)
    assert np.array_equal(x * 2.0 ** 20, np.round(x * 2.0 ** 20)) and (x >= -0.05).all()
    lab = synth.pool_labeled(n, 500)
    assert len(lab) == 500 == len(set(lab.tolist())) and np.array_equal(lab, np.sort(lab))
    assert np.array_equal(lab, synth.pool_labeled(n, 500))


def test_feature_pool_statistics_follow_the_survey_recipe():
    from vatlq import synth
    x = synth.pool_embeddings(3000, kind="clustered").astype(np.float64)
    within = np.linalg.norm(x[0] - x[1])
    across = np.linalg.norm(x[0] - x[30])
    assert 0.5 < within < 0.8 and 15 < across < 22                     # sigma 0.01 noise vs 0.5*ReLU(N(0,1)) centres
    w = synth.pool_embeddings(3000, kind="weak").astype(np.float64)
    assert 2.5 < np.linalg.norm(w[0] - w[30]) / np.linalg.norm(w[0] - w[1]) < 3.5      # the 3:1 control


def test_scale_goldens_are_consistent():
    files = sorted(glob.glob(os.path.join(GOLD, "coreset_scale_*.npz")))
    assert len(files) >= 5
    for f in files:
        z = np.load(f)
        picks = z["picks"]
        assert len(picks) == int(z["k"]) <= int(z["k_full"]) and len(set(picks.tolist())) == len(picks)
        assert 0 <= picks.min() and picks.max() < int(z["n"])
        assert hashlib.sha256(picks.astype("<i8").tobytes()).hexdigest() == str(z["sha256"])
        if bool(z["full_query"]):
            assert int(picks[0]) == int(z["top_unc_idx"][0]) or float(z["labeled_frac"]) > 0    # round 0: first pick = argmax(unc)
