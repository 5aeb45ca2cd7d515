"""Host logic of the Laplacian eigen-solver without a GPU: the subspace iteration of
secedo_b200/csrc/spectral_host.hpp is compiled against a plain-loop backend (tests/spectral_host_check.cpp) and
compared with the oracle (oracle/pyoracle.py: restated laplacian() + LAPACK dsyevd); and the oracle's Laplacian is
pinned by the reference's own known-answer test (tests/test_spectral_clustering.cpp:15-26)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import pyoracle as po
from spectral_cases import noisy_clusters, reference_two_clusters


@pytest.fixture(scope="module")
def host_check(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("spectral") / "spectral_host_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-o", exe, os.path.join(ROOT, "tests", "spectral_host_check.cpp")],
                   check=True)
    return exe


def test_oracle_laplacian_reference_known_answer():
    a = np.array([[0, .5, .2], [.5, 0, .5], [.2, .5, 0]])
    expected = np.array([[1., -0.5976143, -0.28571429], [-0.5976143, 1., -0.5976143], [-0.28571429, -0.5976143, 1.]])
    assert np.abs(po.laplacian(a) - expected).max() < 1e-7  # the reference asserts 1e-3
    z = np.zeros((5, 5))
    assert np.array_equal(po.laplacian(z), np.eye(5))  # zero degree -> 0, not inf (spectral_clustering.cpp:42)


def run_host(exe, tmp_path, a, k, tol=1e-10):
    n = a.shape[0]
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<IId", n, k, tol))
        f.write(np.ascontiguousarray(a, np.float64).tobytes())
    subprocess.run([exe, fin, fout], check=True, capture_output=True)
    raw = open(fout, "rb").read()
    conv, outer = struct.unpack("<II", raw[:8])
    lam = np.frombuffer(raw[8:8 + 8 * k])
    q = np.frombuffer(raw[8 + 8 * k:]).reshape(n, k)
    order = np.argsort(-lam, kind="stable")
    return bool(conv), outer, 1.0 - lam[order], q[:, order]


def check_against_oracle(a, ev, vec, tol):
    from spectral_cases import check_eigenpairs
    check_eigenpairs(a, ev, vec, tol)


CASES = {
    "reference_two_clusters": lambda: (reference_two_clusters(100), 7),
    # every non-trivial eigenvalue of D^-1/2 A D^-1/2 negative (SpectralClustering.OneCluster of the reference)
    "one_cluster_noise": lambda: (1.0 + np.random.default_rng(1243).uniform(-1e-3, 1e-3, (100, 100)), 7),
    "noise_only": lambda: (noisy_clusters(400, 1, 0.0, seed=3), 7),
    "three_weak": lambda: (noisy_clusters(500, 3, 0.2, seed=5), 5),
    "k12": lambda: (noisy_clusters(600, 4, 3.0, seed=9), 12),
    "two_components": lambda: (np.kron(np.eye(2), np.ones((150, 150))) * noisy_clusters(300, 1, 0.0, seed=13), 4),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_subspace_iteration_host_backend(host_check, tmp_path, case):
    a, k = CASES[case]()
    a = np.array(a, dtype=np.float64)
    a = (a + a.T) / 2
    np.fill_diagonal(a, 0)
    conv, outer, ev, vec = run_host(host_check, tmp_path, a, k)
    assert conv and outer < 50
    check_against_oracle(a, ev, vec, 1e-10)


@pytest.mark.parametrize("n", [1, 2, 7, 16, 33, 64])
def test_jacobi_against_lapack(host_check, tmp_path, n):
    """the b x b eigenproblems of the Rayleigh-Ritz steps (cyclic Jacobi on the host) against LAPACK, incl. clustered
    and repeated eigenvalues"""
    rng = np.random.default_rng(n)
    q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    lam = np.sort(np.concatenate([rng.normal(size=n - n // 2), np.full(n // 2, 0.25) + 1e-9 * np.arange(n // 2)]))
    a = (q * lam) @ q.T
    a = (a + a.T) / 2
    fin, fout = str(tmp_path / "j.bin"), str(tmp_path / "jo.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<I", n))
        f.write(a.tobytes())
    subprocess.run([host_check, "--jacobi", fin, fout], check=True)
    raw = np.frombuffer(open(fout, "rb").read())
    w, v = raw[:n], raw[n:].reshape(n, n)
    assert np.abs(w - np.linalg.eigvalsh(a)).max() <= 1e-13 * max(1.0, np.abs(lam).max())
    assert np.abs(v.T @ v - np.eye(n)).max() <= 1e-13
    assert np.abs(a @ v - v * w).max() <= 1e-13 * max(1.0, np.abs(lam).max())


def test_chebyshev_degree_limits(host_check):
    def degree(a, c, top, kth, mmax=100):
        return int(subprocess.run([host_check, "--degree", *map(repr, (a, c, top, kth)), str(mmax)], check=True,
                                  capture_output=True, text=True).stdout)
    # bulk edge: wanted values barely outside the damped interval -> the cap
    assert degree(-0.02, 0.0050, 0.0052, 0.0051) == 100
    # a signal eigenvalue far outside: few steps, so that it cannot swamp the block (growth ratio <= 2e6)
    m = degree(-0.02, 0.005, 0.5, 0.0051)
    assert 2 <= m <= 6
    e, ctr = 0.0125, -0.0075
    growth = np.cosh(m * np.arccosh((0.5 - ctr) / e)) / np.cosh(m * np.arccosh((0.0051 - ctr) / e))
    assert growth <= 2e6 * 1.01
    assert degree(0.0, 0.0, 1.0, 0.5) == 2  # degenerate interval
