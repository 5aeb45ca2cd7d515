"""Similarity matrices for the Laplacian / eigen-solver tests and the comparison with the oracle."""
import numpy as np

from oracle import pyoracle as po


def noisy_clusters(n, n_clusters, signal, seed=0):
    """ADD_MIN-shaped matrix: Gaussian noise plus `signal` for pairs of the same cluster, shifted to >= 0, zero diagonal."""
    rng = np.random.default_rng(seed)
    lab = rng.integers(0, n_clusters, n)
    m = rng.normal(0, 1.0, (n, n))
    m = (m + m.T) / np.sqrt(2)
    m -= signal * (lab[:, None] == lab[None, :])
    np.fill_diagonal(m, 0)
    m = -m
    m += abs(m.min())
    np.fill_diagonal(m, 0)
    return m


def reference_two_clusters(n=100, seed=0):
    """The matrix family of the reference's SpectralClustering.TwoClusters test
    (tests/test_spectral_clustering.cpp:57-97): background 0..5 (sometimes +20), two halves with 100..200."""
    rng = np.random.default_rng(seed)
    a = np.zeros((n, n))
    for i in range(n):
        for j in range(i):
            v = rng.integers(0, 6) + (20 if rng.integers(100, 201) % 5 else 0)
            a[i, j] = a[j, i] = v
    half = n // 2
    for i in range(half):
        for j in range(i):
            if rng.integers(100, 201) % 2:
                a[i, j] = a[j, i] = rng.integers(100, 201)
            else:
                a[i + half, j + half] = a[j + half, i + half] = rng.integers(100, 201)
    return a


def check_eigenpairs(a, ev, vec, tol):
    """ev (k ascending), vec (n, k) against the oracle's full decomposition of laplacian(a): eigenvalues to 100 * tol,
    residuals to 10 * tol, orthonormal, and every vector equal to the oracle's up to sign wherever its eigenvalue is
    separated from its neighbours (error bound residual / gap)."""
    n, k = vec.shape
    lap = po.laplacian(a)
    w, v = np.linalg.eigh(lap)
    assert np.abs(ev - w[:k]).max() <= 100 * tol, (ev, w[:k])
    res = np.linalg.norm(lap @ vec - vec * ev, axis=0)
    assert res.max() <= 10 * tol, res
    assert np.abs(vec.T @ vec - np.eye(k)).max() <= 1e-9
    for i in range(k):
        gap = min(w[i] - w[i - 1] if i else np.inf, w[i + 1] - w[i])
        if gap < 1e-7:
            continue  # (numerically) degenerate: only the invariant subspace is defined
        err = min(np.linalg.norm(vec[:, i] - v[:, i]), np.linalg.norm(vec[:, i] + v[:, i]))
        assert err <= max(1e-9, 20 * tol / gap), (i, err, gap)
    # the subspace as a whole
    proj = v[:, :k].T @ vec
    if k < n and w[k] - w[k - 1] > 1e-7:
        assert np.abs(np.linalg.svd(proj, compute_uv=False) - 1).max() <= max(1e-9, 20 * tol / (w[k] - w[k - 1]))
