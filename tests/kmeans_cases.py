"""Inputs of the K-Means golden cases (tests/golden/kmeans.npz): rebuilt from the counter-based pools, so that the
golden only stores parameters and the reference's outputs."""
import importlib

import numpy as np

CASES = ("clustered", "weak", "iid", "pairs", "one", "all", "dups")


def case_inputs(z, tag):
    synth = importlib.import_module("vatl4pose-wacv2024_b200.synth")
    n, n_lab, seed, k = (int(v) for v in z[f"{tag}_meta"])
    X = synth.pool_embeddings(n, kind=str(z[f"{tag}_kind"]), seed=seed)
    dup = z[f"{tag}_dup"]
    if dup.size:
        X[dup] = X[dup - 100]
    lab = set(synth.pool_labeled(n, n_lab, seed=seed).tolist())
    cand = [i for i in range(n) if i not in lab]
    score = synth.pool_unc(n, seed=seed)[cand]
    w_unc, cw = (float(v) for v in z[f"{tag}_w"])
    return X, cand, score, k, w_unc, cw
