"""GPU parity: heat-map scan (THC + local peaks + coordinates) against the golden fixtures
(reference outputs) and against the oracle on seeded pools; properties at large sizes."""
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # north_star: THC / peak mean within 1e-5 relative in fp32; coordinates bit-exact


def _scan(vatlq, H, ip=None, inx=None, boxes=None, **kw):
    d = "cuda:0"
    r = vatlq.ops.heatmap_scan(torch.from_numpy(H).to(d), ip, inx,
                               None if boxes is None else torch.from_numpy(boxes).to(d), **kw)
    torch.cuda.synchronize()
    return r


def ulp_diff(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


def test_golden_scan(built_lib, gold_scan):
    g = gold_scan
    r = _scan(built_lib, g["H"], g["is_prev"], g["is_next"], g["boxes"])
    assert np.array_equal(r.coords_hm.cpu().numpy(), g["hm_xy"])                      # bit-exact
    k = r.kpts.cpu().numpy()
    assert np.array_equal(k[:, :, 2], g["maxv"])                                      # bit-exact
    d = ulp_diff(k[:, :, :2], g["img_xy"])
    assert d.max() <= 1 and (d > 0).mean() < 1e-3, (d.max(), (d > 0).mean())          # cv2 LU vs closed form
    assert np.allclose(r.thc.cpu().numpy(), g["thc"], rtol=RTOL, atol=0)
    assert np.allclose(r.peak_mean.cpu().numpy(), g["peak"], rtol=RTOL, atol=0, equal_nan=True)
    assert np.isnan(r.peak_mean.cpu().numpy()[11]) and r.peak_cnt.cpu().numpy()[11] == 0


def test_reference_signature_shims(built_lib, gold_scan):
    v = built_lib
    g = gold_scan
    H = g["H"]
    preds, maxvals = v.heatmap_to_coord_simple(H[0], g["boxes"][0].tolist())
    assert preds.shape == (17, 2) and maxvals.shape == (17, 1)
    assert ulp_diff(preds, g["img_xy"][0]).max() <= 1 and np.array_equal(maxvals[:, 0], g["maxv"][0])
    assert np.isclose(v.localpeak_mean(H[0]), g["peak"][0], rtol=RTOL)
    from oracle import vatl_oracle as O
    assert np.isclose(v.compute_thc(H[1], H[0]), float(O.thc_pair(H[1], H[0])), rtol=RTOL)


@pytest.mark.parametrize("n,seed,mean_len", [(256, 0, 30.0), (67, 5, 3.0), (1, 6, 30.0), (2, 7, 30.0)])
def test_scan_vs_oracle(built_lib, n, seed, mean_len):
    """Config 1 (256 frames x 17 x 64x48) and ragged small pools."""
    from oracle import vatl_oracle as O
    synth = built_lib.synth
    ids, ip, inx = synth.track_flags(n, np.random.default_rng(seed), mean_len)
    H = synth.heatmaps(n, seed=seed, track_ids=ids)
    boxes = synth.boxes_xyxy(n, seed)
    r = _scan(built_lib, H, ip, inx, boxes)
    hm = r.coords_hm.cpu().numpy()
    k = r.kpts.cpu().numpy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(n):
            xy, mv = O.heatmap_coords(H[i])
            assert np.array_equal(hm[i], xy) and np.array_equal(k[i, :, 2], mv[:, 0])
            if i % 8 == 0:
                c, _ = O.heatmap_to_coord(H[i], boxes[i].tolist())
                assert ulp_diff(k[i, :, :2], c).max() <= 1
                assert np.isclose(r.peak_mean[i].item(), O.localpeak_mean(H[i]), rtol=RTOL)
    assert np.allclose(r.thc.cpu().numpy(), O.thc_pool(H, ip, inx), rtol=RTOL, atol=0)


def test_empty_pool(built_lib):
    r = built_lib.ops.heatmap_scan(torch.zeros((0, 17, 64, 48), device="cuda:0"))
    assert r.thc.numel() == 0


def test_halo_and_chunking_equal_whole_pool(built_lib):
    """A pool scanned as shards with halo frames / in chunks gives the bits of one scan."""
    v = built_lib
    n = 300
    ids, ip, inx = v.synth.track_flags(n, np.random.default_rng(3), 12.0)
    H = torch.from_numpy(v.synth.heatmaps(n, seed=3, track_ids=ids)).cuda()
    bb = torch.from_numpy(v.synth.boxes_xyxy(n, 3)).cuda()
    whole = v.ops.heatmap_scan(H, ip, inx, bb)
    for cut in (1, 100, 151, 299):
        a = v.ops.heatmap_scan(H[:cut], ip[:cut], inx[:cut], bb[:cut], halo_next=H[cut])
        b = v.ops.heatmap_scan(H[cut:], ip[cut:], inx[cut:], bb[cut:], halo_prev=H[cut - 1])
        assert torch.equal(torch.cat([a.thc, b.thc]), whole.thc)
        assert torch.equal(torch.cat([a.kpts, b.kpts]), whole.kpts)
        assert torch.equal(torch.cat([a.peak_sum, b.peak_sum]), whole.peak_sum)
    qp = v.QueryPass(n, "cuda:0", uncertainty="THC")
    qp.score_pool(H, bb, torch.from_numpy(ip), torch.from_numpy(inx), chunk=64)
    assert torch.equal(qp.thc, whole.thc) and torch.equal(qp.kpts, whole.kpts)
    assert torch.equal(qp.peak_mean, whole.peak_mean, ) or torch.allclose(qp.peak_mean, whole.peak_mean, equal_nan=True)


def test_thc3_strict_dropin(built_lib):
    from oracle import vatl_oracle as O
    v = built_lib
    n = 40
    rng = np.random.default_rng(9)
    cur, prv, nxt = (v.synth.heatmaps(n, seed=s) for s in (20, 21, 22))
    ip, inx = rng.integers(0, 2, n).astype(np.uint8), rng.integers(0, 2, n).astype(np.uint8)
    out = v.ops.thc3(torch.from_numpy(cur).cuda(), torch.from_numpy(prv).cuda(), torch.from_numpy(nxt).cuda(), ip, inx)
    ref = [O.thc_item(cur[i], prv[i], nxt[i], bool(ip[i]), bool(inx[i])) for i in range(n)]
    assert np.allclose(out.cpu().numpy(), ref, rtol=RTOL, atol=0)


def test_generic_shape_path(built_lib):
    """Odd map shapes take the generic kernel: the 4x10 known-answer map of local_peak.py:25-31."""
    from oracle import vatl_oracle as O
    kat = np.array([[0, 0, 0, 0, 0, 0, 0, 4, 0, 0], [0, 0, 0, 1, 1, 0, 0, 0, 0, 0],
                    [0, 0, 0, 0, 3, 2, 0, 0, 0, 0], [0, 0, 0, 0, 2, 2, 0, 0, 0, 0]], np.float32)
    H = np.stack([kat, kat.T.copy().reshape(4, 10) * 0.5, -kat])[None].repeat(3, 0)     # (3,3,4,10)
    H[1] *= 0.7
    r = _scan(built_lib, H, np.array([0, 1, 1], np.uint8), np.array([1, 1, 0], np.uint8))
    assert r.peak_sum.cpu().numpy()[0] == pytest.approx(sum(O.localpeak_values(m).sum() for m in H[0]))
    assert r.peak_cnt.cpu().tolist()[0] == sum(O.localpeak_values(m).size for m in H[0])
    assert np.allclose(r.thc.cpu().numpy(), O.thc_pool(H, [0, 1, 1], [1, 1, 0]), rtol=RTOL)
    for i in range(3):
        assert np.array_equal(r.coords_hm.cpu().numpy()[i], O.heatmap_coords(H[i])[0])


def test_large_pool_properties(built_lib):
    """Full-size behaviour through size-independent properties (20k frames = 4.2 GB):
    THC is symmetric in time (reversing the pool reverses the scores), chunk-invariant, and the
    peak values equal the torch max of every map."""
    v = built_lib
    n = 20000
    H, ip, inx, bb = v.synth.device_pool(n, "cuda:0", seed=4)
    a = v.ops.heatmap_scan(H, ip, inx, bb)
    assert torch.equal(a.kpts[:, :, 2], H.flatten(2).max(dim=2).values)
    flat_arg = H.flatten(2).argmax(dim=2)
    base = torch.stack([(flat_arg % 48).float(), (flat_arg // 48).float()], dim=2)
    assert (a.coords_hm - base).abs().max().item() <= 0.25
    Hr = H.flip(0).contiguous()
    b = v.ops.heatmap_scan(Hr, inx.flip(0).contiguous(), ip.flip(0).contiguous(), bb.flip(0).contiguous())
    assert torch.allclose(b.thc.flip(0), a.thc, rtol=1e-6, atol=0)
    assert torch.equal(b.peak_cnt.flip(0), a.peak_cnt)
    pair = (H[1:] - H[:-1]).abs().flatten(1).double().sum(1) / 17
    both = (ip[1:-1] == 1) & (inx[1:-1] == 1)
    expect = (pair[:-1] + pair[1:])[both]
    assert torch.allclose(a.thc[1:-1][both].double(), expect, rtol=1e-5, atol=0)
