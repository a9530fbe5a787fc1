"""GPU parity: WPU (hybrid feature + auto-encoder + MSE) and the score fusion."""
import numpy as np
import pytest
import torch

from conftest import ae_weights_from_gold

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def test_golden_wpu(built_lib, gold_wpu):
    v, g = built_lib, gold_wpu
    w, ind, z = v.ops.pack_ae_weights(ae_weights_from_gold(g), "cuda:0")
    k = torch.from_numpy(g["kpts"]).cuda()
    b = torch.from_numpy(g["boxes"]).cuda()
    w42, feat = v.ops.wpu(k, b, w, ind, z, drop_ears=False, return_features=True)
    w38 = v.ops.wpu(k, b, w, ind, z, drop_ears=True)
    # the fp64 feature stage reproduces the reference's float32 inputs bit for bit
    ref_feat = g["feat"].astype(np.float32)
    same = (feat.cpu().numpy() == ref_feat)
    assert same.mean() > 0.999, same.mean()
    assert np.allclose(feat.cpu().numpy(), ref_feat, rtol=1e-6, atol=1e-7)
    assert np.allclose(w42.cpu().numpy(), g["wpu42"], rtol=RTOL, atol=0)
    assert np.allclose(w38.cpu().numpy(), g["wpu38"], rtol=RTOL, atol=0)


def test_config2_wpu_10k_poses(built_lib):
    """Config 2: 10 k synthetic poses; oracle on a sample, finiteness and range on all."""
    from oracle import vatl_oracle as O
    v = built_lib
    kp, bb = v.synth.poses(10000, seed=1)
    W = v.synth.ae_weights(42, 4, seed=318)
    w, ind, z = v.ops.pack_ae_weights(W, "cuda:0")
    out = v.ops.wpu(torch.from_numpy(kp).cuda(), torch.from_numpy(bb).cuda(), w, ind, z).cpu().numpy()
    assert np.isfinite(out).all() and (out >= 0).all()
    ae = O.make_autoencoder(W)
    for i in range(0, 10000, 97):
        ref = O.wpu_item(ae, bb[i].tolist(), kp[i].reshape(-1).astype(np.float64))
        assert np.isclose(out[i], ref, rtol=RTOL, atol=0), i
    ae2 = v.WholeBodyAE(z_dim=4)
    s = ae2.unnaturalness(kp[:8], bb[:8]).cpu().numpy()
    with torch.no_grad():
        o = O.make_autoencoder([(m.weight.detach().numpy(), m.bias.detach().numpy())
                                for m in list(ae2.encoder) + list(ae2.decoder) if isinstance(m, torch.nn.Linear)])
    assert np.isclose(s[3], O.wpu_item(o, bb[3].tolist(), kp[3].reshape(-1).astype(np.float64)), rtol=RTOL)


def test_wpu_status_flags(built_lib):
    v = built_lib
    kp, bb = v.synth.poses(4, seed=2)
    bb[1, 3] = bb[1, 1] - 5.0          # height <= 0
    kp[2, :, 2] = 0.0                  # sum(scores) <= 0
    w, ind, z = v.ops.pack_ae_weights(v.synth.ae_weights(), "cuda:0")
    with pytest.raises(AssertionError, match="height"):
        v.ops.wpu(torch.from_numpy(kp[:2]).cuda(), torch.from_numpy(bb[:2]).cuda(), w, ind, z)
    with pytest.raises(AssertionError, match="visible"):
        v.ops.wpu(torch.from_numpy(kp[2:]).cuda(), torch.from_numpy(bb[2:]).cuda(), w, ind, z)
    out = v.ops.wpu(torch.from_numpy(kp).cuda(), torch.from_numpy(bb).cuda(), w, ind, z, check_status=False)
    assert torch.isnan(out[1]) and torch.isnan(out[2]) and torch.isfinite(out[0]) and torch.isfinite(out[3])


def test_golden_fusion_bit_exact(built_lib, gold_fuse):
    v, g = built_lib, gold_fuse
    # fp32 inputs as the device path holds them; the reference formulas evaluated on the same values
    from oracle import vatl_oracle as O
    t32, w32 = g["thc"].astype(np.float32), g["wpu"].astype(np.float32)
    n = t32.size
    unl = np.ones(n, np.uint8); unl[[3, 17, 40]] = 0
    sel = unl.astype(bool)
    for mode in ("const", "increase", "decrease"):
        ref = O.fuse_scores(t32[sel].astype(np.float64), w32[sel].astype(np.float64), mode, 0.15)
        out = v.ops.fuse_scores(torch.from_numpy(t32).cuda(), torch.from_numpy(w32).cuda(), torch.from_numpy(unl).cuda(),
                                mode, 0.15).cpu().numpy()
        assert np.array_equal(out[sel], ref), mode          # IEEE double ops in the reference's order
        assert (out[~sel] == 0).all()
    single = v.ops.fuse_scores(torch.from_numpy(t32).cuda(), None, None).cpu().numpy()
    assert np.array_equal(single, O.fuse_scores(t32.astype(np.float64)))


def test_oks_matches_reference_golden(built_lib):
    """vatlq_oks against al_metric.compute_OKS outputs (fp64; exp differs from libm by <= 1 ulp)."""
    import torch
    from conftest import load_golden
    z = load_golden("oks.npz")
    got = built_lib.ops.oks(torch.from_numpy(z["kpts"]).cuda(), torch.from_numpy(z["gt"]).cuda(),
                            torch.from_numpy(z["boxes"]).cuda()).cpu().numpy()
    assert np.allclose(got, z["oks"], rtol=1e-12, atol=1e-300)
    assert got[7] == 1.0
