"""Laplacian and its leading eigenpairs on the GPU (spectral_clustering.cpp:33-52, :127-138)
through the C ABI, against the oracle (restated laplacian() + LAPACK dsyevd, which is what arma::eig_sym calls).
Floating point: eigenvalues within 1e-8, residuals |L v - lambda v| within 1e-9, vectors up to sign within
residual / gap (spectral_cases.check_eigenpairs)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from secedo_b200 import api
from secedo_b200.synth import SynthConfig, make_pileup
from spectral_cases import check_eigenpairs, noisy_clusters, reference_two_clusters

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.mark.parametrize("kernel", ["tma", "registers"])
@pytest.mark.parametrize("width", [8, 16, 32, 64])
@pytest.mark.parametrize("n", [70, 257, 1000, 2100])
def test_block_product_kernel(gpu_ctx, n, width, kernel, monkeypatch):
    """alpha M X + beta X + gamma W of the solver's only O(N^2) kernel (split over K, fixed summation order): the
    TMA-streamed kernel (widths <= 32) and the register-prefetch kernel (width 64, or all widths on request)"""
    if kernel == "registers":
        monkeypatch.setenv("SECEDO_B200_SPECTRAL_NO_TMA", "1")
    rng = np.random.default_rng(n + width)
    m = rng.normal(size=(n, n))
    m = m + m.T
    x, w = rng.normal(size=(n, width)), rng.normal(size=(n, width))
    for alpha, beta, gamma in ((1.0, 0.0, 0.0), (0.7, -0.3, 0.0), (2.5, 0.1, -1.25)):
        out = api.spectral_matvec(m, x, w, alpha, beta, gamma, ctx=gpu_ctx)
        ref = alpha * (m @ x) + beta * x + gamma * w
        assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max() * np.sqrt(n)
    a = api.spectral_matvec(m, x, w, 2.5, 0.1, -1.25, ctx=gpu_ctx)
    assert np.array_equal(a, out), "the block product must be deterministic"


def test_laplacian_matches_oracle(gpu_ctx):
    a = np.array([[0, .5, .2], [.5, 0, .5], [.2, .5, 0]])  # tests/test_spectral_clustering.cpp:15-26
    assert np.abs(api.laplacian(a, gpu_ctx) - po.laplacian(a)).max() <= 1e-15
    a = noisy_clusters(333, 3, 1.0, seed=2)
    a[17, :] = 0
    a[:, 17] = 0  # an isolated cell: degree 0 -> scale 0 (spectral_clustering.cpp:42)
    lap, ref = api.laplacian(a, gpu_ctx), po.laplacian(a)
    assert np.abs(lap - ref).max() <= 1e-14
    assert lap[17, 17] == 1.0 and np.array_equal(np.diag(lap), np.ones(333))


CASES = {
    "reference_two_clusters": lambda: (reference_two_clusters(100), 7),
    "reference_three_clusters_99": lambda: (reference_two_clusters(99, seed=4), 3),
    "one_cluster_noise": lambda: (1.0 + np.random.default_rng(1243).uniform(-1e-3, 1e-3, (100, 100)), 7),
    "noise_700": lambda: (noisy_clusters(700, 1, 0.0, seed=3), 7),
    "weak_clusters_1500": lambda: (noisy_clusters(1500, 3, 0.15, seed=5), 7),
    "strong_clusters_2000_k12": lambda: (noisy_clusters(2000, 4, 5.0, seed=9), 12),
    "isolated_cells": lambda: (noisy_clusters(300, 2, 2.0, seed=11), 5),
    "two_components": lambda: (np.kron(np.eye(2), np.ones((150, 150))) * noisy_clusters(300, 1, 0.0, seed=13), 4),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_spectral_embedding_vs_oracle(gpu_ctx, name):
    a, k = CASES[name]()
    a = np.array(a, dtype=np.float64)
    a = (a + a.T) / 2
    np.fill_diagonal(a, 0)
    if name == "isolated_cells":
        a[[5, 77], :] = 0
        a[:, [5, 77]] = 0
    ev, vec, st = api.spectral_embedding(a, k, TOL, ctx=gpu_ctx, return_stats=True)
    assert st["max_residual"] <= TOL and st["outer_iterations"] < 60 and st["matvec_launches"] > 0
    assert ev[0] == 0.0 or abs(ev[0]) < 1e-12
    check_eigenpairs(a, ev, vec, TOL)
    # first eigenvector: sqrt(degree), the one the reference's k-means keeps (spectral_clustering.cpp:170-171)
    d = np.sqrt(a.sum(1))
    if name != "two_components":
        assert np.abs(vec[:, 0] - d / np.linalg.norm(d)).max() <= 1e-12


def test_spectral_edge_cases(gpu_ctx):
    ev, vec = api.spectral_embedding(np.zeros((99, 99)), 5, ctx=gpu_ctx)  # SpectralClustering.AllZero of the reference
    assert np.array_equal(ev, np.ones(5)) and np.array_equal(vec, np.eye(99)[:, :5])
    a = noisy_clusters(100, 2, 1.0)
    ev, vec = api.spectral_embedding(a, 1, ctx=gpu_ctx)
    assert ev[0] == 0.0 and np.abs(vec[:, 0] - np.sqrt(a.sum(1) / a.sum())).max() < 1e-15
    with pytest.raises(api.SgpuError):
        api.spectral_embedding(-a, 3, ctx=gpu_ctx)  # negative degrees: sqrt of a negative number in the reference
    with pytest.raises(api.SgpuError):
        api.spectral_embedding(a[:20, :20], 7, ctx=gpu_ctx)  # fewer cells than twice the block
    with pytest.raises(api.SgpuError):
        api.spectral_embedding(a, 40, ctx=gpu_ctx)


def test_finalize_spectral_on_device_matrix(gpu_ctx):
    """pileup -> counts -> epilogue -> Laplacian -> eigenpairs without the matrix leaving HBM equals the oracle's
    decomposition of the matrix that sgpu_similarity_finalize returns"""
    cfg = SynthConfig(n_cells=200, coverage=0.4, n_loci=3000, n_chr=2, n_clones=2, p_multi=0.1, p_mate=0.05, theta=0.01, seed=21)
    ident = np.arange(cfg.n_cells, dtype=np.uint32)
    f, _ = api.Filter(0.01, 4, gpu_ctx).filter(make_pileup(cfg), ident, "", 1)
    c = api.Counts(gpu_ctx, cfg.n_cells)
    c.accumulate(f, 1000, ident, 0.01, 0.5, 0.01, 8, "gemm")
    m = c.finalize(1000, 0.01, 0.5, 0.01, "ADD_MIN")
    ev, vec, st, m2 = c.finalize_spectral(1000, 0.01, 0.5, 0.01, "ADD_MIN", k=7, tol=TOL, want_matrix=True)
    assert np.array_equal(m, m2)
    check_eigenpairs(m, ev, vec, TOL)
    ev3, vec3, _, none = c.finalize_spectral(1000, 0.01, 0.5, 0.01, "ADD_MIN", k=7, tol=TOL)
    assert none is None and np.array_equal(ev3, ev) and np.array_equal(vec3, vec)
    c.free()


def test_full_size_residuals(gpu_ctx):
    """8 000 cells (cfg3): 7 eigenpairs of the device-resident matrix; checked through size-independent properties
    (residuals, orthonormality, closed-form first eigenvector) - the full LAPACK decomposition takes minutes"""
    n = 8000
    dev = gpu_ctx.synth_pileup(n, 0.5, 1, 2048, n_clones=4, theta=0.001, p_multi=0.005, p_mate=0.01, seed=5)
    ident = np.arange(n, dtype=np.uint32)
    fdev, _ = api.Filter(0.001, 4, gpu_ctx).filter_device(dev, ident)
    c = api.Counts(gpu_ctx, n)
    c.accumulate(fdev, 1000, ident, 0.01, 0.15, 0.001, 8, "gemm")
    ev, vec, st, m = c.finalize_spectral(1000, 0.01, 0.15, 0.001, "ADD_MIN", k=7, tol=TOL, want_matrix=True)
    c.free()
    d = m.sum(1)
    s = 1 / np.sqrt(d)
    bv = s[:, None] * (m @ (s[:, None] * vec))  # B vec; L vec = vec - B vec
    res = np.linalg.norm(vec - bv - vec * ev, axis=0)
    assert res.max() <= 1e-9, res
    assert np.abs(vec.T @ vec - np.eye(7)).max() <= 1e-9
    assert np.all(np.diff(ev) >= -1e-12) and ev[0] == 0.0
    assert np.abs(vec[:, 0] - np.sqrt(d / d.sum())).max() <= 1e-12
    print("spectral stats at 8000 cells:", st)
