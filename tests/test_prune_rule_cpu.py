"""CPU: the pruning RULE of the core-set passes (coreset.cu "pruning": segments of consecutive rows, anchor +
radius, triangle inequality with a margin) restated in numpy and checked against brute force: a segment the
rule flags must never contain a row that a centre of the pass would move.  This checks the mathematics and
the margin on adversarial inputs (duplicates, zero rows, collinear points, huge and tiny norms, ragged and
singleton tracks); tests/test_gpu_prune.py checks that the kernels implement it bit for bit."""
import zlib

import numpy as np
import pytest

SEG_MAX = 64


def sk_dist(X, xx, c):
    """The canonical distance the passes use: sqrt(max(0, (-2 x.c + |x|^2) + |c|^2)) in float64."""
    return np.sqrt(np.maximum((-2.0 * (X @ X[c]) + xx) + xx[c], 0.0))


def segments(X):
    """segment_kernel: cut where the consecutive-row distance exceeds twice its mean, and every SEG_MAX rows."""
    n = X.shape[0]
    cd = np.full(n, np.inf)
    cd[1:] = np.sqrt(((X[1:] - X[:-1]) ** 2).sum(1))
    fin = np.isfinite(cd)
    tau = 2.0 * cd[fin].sum() / fin.sum() if fin.any() else np.inf
    starts, rs = [], 0
    for i in range(n):
        cut = i == 0 or cd[i] > tau
        if cut:
            rs = i
        else:
            cut = (i - rs) % SEG_MAX == 0
        if cut:
            starts.append(i)
    return np.array(starts + [n])


def run_case(X32, labeled, k, rng, group=8):
    X = X32.astype(np.float64)
    n = X.shape[0]
    xx = np.einsum("ij,ij->i", X, X)
    st = segments(X)
    nseg = len(st) - 1
    assert np.diff(st).max() <= SEG_MAX and np.diff(st).min() >= 1
    R = np.array([np.sqrt(((X[st[s]:st[s + 1]] - X[st[s]]) ** 2).sum(1)).max() for s in range(nseg)])
    md = np.full(n, np.inf)
    for c in labeled:
        md = np.minimum(md, sk_dist(X, xx, c))
    flagged_rows = total_rows = 0
    picks = []
    while len(picks) < k:
        # a "round": the next `group` greedy picks decided sequentially on a scratch copy, applied as one pass
        scratch, cs = md.copy(), []
        for _ in range(min(group, k - len(picks))):
            c = int(np.argmax(scratch)) if np.isfinite(scratch).all() else int(rng.integers(n))
            cs.append(c)
            scratch = np.minimum(scratch, sk_dist(X, xx, c))
        picks += cs
        cn = np.sqrt(max(xx[c] for c in cs))
        new = md.copy()
        for c in cs:
            new = np.minimum(new, sk_dist(X, xx, c))
        moved = new != md
        for s in range(nseg):
            a, e = st[s], st[s + 1]
            M = md[a:e].max()
            dmin = min(sk_dist(X[[a] + cs], xx[[a] + cs], j + 1)[0] for j in range(len(cs)))
            margin = 1e-6 * (1.0 + np.sqrt(xx[a]) + cn)
            flag = np.isfinite(M) and (dmin - R[s] - M >= margin)
            total_rows += e - a
            if flag:
                flagged_rows += e - a
                assert not moved[a:e].any(), (s, a, e, M, dmin, R[s])
                assert not any(a <= c < e for c in cs)          # a picked row is never pruned
        md = new
    return flagged_rows / max(total_rows, 1)


def clustered(n, d, rng, sizes=(1, 90), scale=0.5, noise=0.01):
    out, cen = [], None
    while sum(len(b) for b in out) < n:
        m = int(rng.integers(sizes[0], sizes[1] + 1))
        cen = np.maximum(rng.standard_normal(d), 0) * scale
        out.append(cen + rng.normal(0, noise, (m, d)))
    return np.concatenate(out)[:n].astype(np.float32)


@pytest.mark.parametrize("name", ["tracks", "iid", "duplicates_and_zero_rows", "collinear", "huge_norms", "tiny_norms",
                                  "labelled"])
def test_flagged_segments_never_move(name):
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    d, labeled, k = 64, [], 96
    if name == "tracks":
        X = clustered(1500, d, rng)
    elif name == "iid":
        X = np.maximum(rng.standard_normal((800, d)), 0).astype(np.float32)
    elif name == "duplicates_and_zero_rows":
        X = clustered(900, d, rng, sizes=(5, 40))
        X[100:140] = X[100]            # a track of identical frames
        X[300:310] = 0                 # zero rows
        X[500] = X[20]                 # a far duplicate
    elif name == "collinear":
        t = np.sort(rng.uniform(0, 50, 700))[:, None]
        X = (t * np.ones((1, d)) / np.sqrt(d)).astype(np.float32)     # points on a line: the bound is tight
    elif name == "huge_norms":
        X = (clustered(700, d, rng) * 3.0e4).astype(np.float32)
    elif name == "tiny_norms":
        X = (clustered(700, d, rng) * 1.0e-4).astype(np.float32)
    else:
        X = clustered(1200, d, rng, sizes=(10, 50))
        labeled = sorted(rng.choice(1200, 150, replace=False).tolist())
    frac = run_case(X, labeled, k, rng)
    if name in ("tracks", "labelled", "duplicates_and_zero_rows"):
        assert frac > 0.2, frac        # the rule actually prunes on track-like pools
