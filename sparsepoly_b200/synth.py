"""Deterministic synthetic workloads of SURVEY.md 8d / BASELINE.json (host-side numpy)."""
import numpy as np
import scipy.sparse as sp


def uniform_sparse(n, d, nnz_row, seed):
    """CSR: every row draws nnz_row column ids i.i.d. uniform on [0,d), values N(0,1), duplicates
    summed, indices sorted (C1-C4 inputs)."""
    rng = np.random.RandomState(seed)
    cols = rng.randint(0, d, size=(n, nnz_row)).astype(np.int32)
    vals = rng.randn(n, nnz_row)
    indptr = np.arange(0, n * nnz_row + 1, nnz_row, dtype=np.int64)
    X = sp.csr_matrix((vals.ravel(), cols.ravel(), indptr), shape=(n, d))
    X.sum_duplicates()
    X.sort_indices()
    return X


def criteo_like(n, d, seed, n_numeric=13, n_categorical=26, zipf_a=1.1):
    """CSR with exactly n_numeric + n_categorical nonzeros per row (C5): numeric field f is column f
    with a U(0,1) value; categorical field g owns an equal contiguous slice of the remaining
    columns and activates one id drawn Zipf(zipf_a) inside the slice with value 1."""
    rng = np.random.RandomState(seed)
    r = n_numeric + n_categorical
    cols = np.empty((n, r), dtype=np.int32)
    vals = np.empty((n, r), dtype=np.float64)
    cols[:, :n_numeric] = np.arange(n_numeric, dtype=np.int32)
    vals[:, :n_numeric] = rng.rand(n, n_numeric)
    width = (d - n_numeric) // n_categorical
    if width < 1:
        raise ValueError("d too small for the Criteo-shaped layout")
    cdf = np.cumsum(np.arange(1, width + 1, dtype=np.float64) ** (-zipf_a))
    cdf /= cdf[-1]
    for g in range(n_categorical):           # truncated Zipf(a) over the slice, inverse-CDF sampling
        ids = np.searchsorted(cdf, rng.rand(n), side="left").astype(np.int32)
        cols[:, n_numeric + g] = n_numeric + g * width + np.minimum(ids, width - 1)
    vals[:, n_numeric:] = 1.0
    indptr = np.arange(0, n * r + 1, r, dtype=np.int64)
    X = sp.csr_matrix((vals.ravel(), cols.ravel(), indptr), shape=(n, d))
    X.sort_indices()
    return X


def planted_P(d, k_true, seed, frac_features=0.1, scale=0.3):
    """P* [k_true, d]: nonzero rows confined to a random 10% of the features."""
    rng = np.random.RandomState(seed)
    active = rng.choice(d, size=max(1, int(frac_features * d)), replace=False)
    P = np.zeros((k_true, d))
    P[:, active] = scale * rng.randn(k_true, active.size)
    return P


def targets_from_scores(scores, seed, classification, noise=0.1):
    rng = np.random.RandomState(seed)
    y = scores + noise * np.std(scores) * rng.randn(scores.shape[0])
    if classification:
        return np.where(y > np.median(y), 1.0, -1.0)
    return y


def planted_fm_targets(X, model_seed, noise_seed, k_true=4, positive_frac=0.25, classification=True):
    """Targets of a planted sparse degree-2 FM: linear + pairwise signal confined to a random
    10 % of the features, plus noise; classification thresholds at the (1 - positive_frac) quantile.  Host
    side only (scipy), so that both bench arms see the same data."""
    n, d = X.shape
    rng = np.random.RandomState(model_seed)
    active = rng.choice(d, size=max(2, d // 10), replace=False)
    Pt = np.zeros((d, k_true))
    Pt[active] = 0.3 * rng.randn(active.size, k_true)
    wt = np.zeros(d)
    wt[active] = 0.3 * rng.randn(active.size)
    X2 = X.copy()
    X2.data = X2.data ** 2
    XP = X @ Pt
    s = X @ wt + 0.5 * (XP ** 2 - X2 @ (Pt ** 2)).sum(1)
    s = np.asarray(s).ravel()
    nrng = np.random.RandomState(noise_seed)
    s = s + 0.1 * np.std(s) * nrng.randn(n)
    if classification:
        return np.where(s > np.quantile(s, 1.0 - positive_frac), 1.0, -1.0)
    return s / np.std(s)
