"""CPU restatements of two index / ordering schemes the kernels rely on, checked against the library semantics they
must reproduce (the kernels themselves are checked bit for bit on the GPU: test_gpu_next.py, test_gpu_coreset.py).

* MPE / Margin register path (csrc/extras.cu, peak_unc_reg_kernel): lane l holds the 11-wide row maxima of rows 2l and
  2l+1; the 11-tall column maxima come from the pair maxima of lanes l-2 .. l+2, the second row of lane l-3 and the
  first row of lane l+3, source lanes clamped to 0 .. 31.  Must equal skimage's maximum_filter(size=11,
  mode='nearest') (peak_local_max, ActiveLearning.py:766,784).
* planner arg-max (csrc/coreset.cu, score_key / warp_argmax_key): scores are reduced as order-preserving 64-bit integer
  keys, ties by the lowest row id.  Must equal np.argmax over the fp64 scores (ActiveLearning.py:822,828,834).
"""
import numpy as np


def test_lane_shuffle_column_maxima_equal_maximum_filter():
    from scipy.ndimage import maximum_filter
    rng = np.random.default_rng(0)
    lanes = np.arange(32)
    for trial in range(12):
        img = rng.normal(0, 1, (64, 48)).astype(np.float32)
        if trial % 3 == 0:
            img = np.round(img)                              # plateaus / ties
        if trial == 5:
            img[:] = 0.25                                    # constant map
        pad = np.pad(img, ((0, 0), (5, 5)), mode="edge")
        rowmax = np.max(np.stack([pad[:, i:i + 48] for i in range(11)]), axis=0)
        a, b = rowmax[0::2], rowmax[1::2]                    # lane l: rows 2l, 2l+1
        src = lambda d: np.clip(lanes + d, 0, 31)
        m = np.maximum(a, b)
        m5 = np.maximum.reduce([m, m[src(-1)], m[src(-2)], m[src(1)], m[src(2)]])
        out = np.empty_like(img)
        out[0::2] = np.maximum(m5, b[src(-3)])
        out[1::2] = np.maximum(m5, a[src(3)])
        assert np.array_equal(out, maximum_filter(img, size=11, mode="nearest"))


def _score_key(s):
    s = np.where(s == 0.0, 0.0, s)                           # -0.0 and +0.0 share a key
    bits = s.view(np.uint64)
    neg = (bits >> np.uint64(63)).astype(bool)
    return np.where(neg, ~bits, bits | np.uint64(1 << 63))


def test_integer_score_keys_reproduce_argmax_with_lowest_index_ties():
    rng = np.random.default_rng(1)
    special = np.array([0.0, -0.0, np.inf, -np.inf, 5e-324, -5e-324, 1e308, -1e308, 1.0, np.nextafter(1.0, 2.0)])
    for trial in range(50):
        s = np.round(rng.normal(0, 1, 300), 1)               # many exact ties
        s[rng.choice(300, len(special), replace=False)] = special
        if trial % 2:
            s[s == np.inf] = 0.0                             # let the finite maximum decide
        key = _score_key(s.astype(np.float64))
        idx = np.arange(300)
        winner = idx[key == key.max()].min()                 # largest key, lowest row id on equal keys
        assert winner == int(np.argmax(s))
        canon = np.where(s == 0, 0.0, s)                     # the key order is the fp64 order (signed zeros merged)
        assert np.array_equal(np.sort(canon), canon[np.argsort(key, kind="stable")])
