"""expectation_maximization on the GPU (expectation_maximization.cpp:131-160) through the C ABI:
the reference's own known-answer tests, the golden vectors of the compiled reference and the oracle on synthetic
pileups. The per-cell sums are formed in 2^-32 fixed point instead of the reference's sequential fp64 order, so the
probabilities are compared within 1e-6 (absolute; they are probabilities) and the iteration counts must agree."""
import numpy as np
import pytest

from oracle import pyoracle as po
from secedo_b200 import api
from secedo_b200.pileup import NO_POS
from secedo_b200.synth import SynthConfig, make_pileup
from test_oracle_em import REFERENCE_TESTS, THETA, golden_em_cases, reference_test_input, repeated_locus

pytestmark = pytest.mark.gpu
TOL = 1e-6


@pytest.mark.parametrize("name", sorted(REFERENCE_TESTS))
def test_reference_known_answers(gpu_ctx, name):
    p, m, start, check = reference_test_input(name)
    prob, st = api.expectation_maximization(p, m, 1, THETA, start, ctx=gpu_ctx, return_stats=True)
    assert check(prob), prob
    want, it = po.expectation_maximization(p, m, THETA, start)
    assert np.abs(prob - want).max() <= TOL and st["iterations"] == it


def test_golden_vectors(gpu_ctx):
    for name, p, m, theta, start, final, iterations in golden_em_cases():
        prob, st = api.expectation_maximization(p, m, 8, theta, start, ctx=gpu_ctx, return_stats=True)
        assert np.abs(prob - final).max() <= TOL, name
        assert st["iterations"] == iterations, name
        again = api.expectation_maximization(p, m, 8, theta, start, ctx=gpu_ctx)
        assert np.array_equal(prob, again), "fixed-point sums must make the result bit-reproducible"


@pytest.mark.parametrize("n_cells,coverage", [(300, 0.3), (2000, 0.1), (7000, 0.05)])
def test_against_oracle_on_filtered_device_pileup(gpu_ctx, n_cells, coverage):
    """filter on the device, then EM on the device-resident result (the way divide_cluster chains them); the
    intermediate states after 1 and 2 iterations are compared as well"""
    cfg = SynthConfig(n_cells=n_cells, coverage=coverage, n_loci=1500, n_chr=2, n_clones=2, p_multi=0.05, p_mate=0.02, theta=0.01,
                      seed=n_cells)
    ident = np.arange(n_cells, dtype=np.uint32)
    fdev, _ = api.Filter(0.01, 4, gpu_ctx).filter_device(make_pileup(cfg), ident)
    f = fdev.download()
    rng = np.random.default_rng(n_cells)
    start = np.clip((np.arange(n_cells) >= n_cells // 2) * 0.5 + 0.25 + rng.normal(0, 0.1, n_cells), 0.02, 0.98)
    want, it = po.expectation_maximization(f, ident, 0.01, start)
    prob, st = api.expectation_maximization(fdev, ident, 8, 0.01, start, ctx=gpu_ctx, return_stats=True)
    assert st["iterations"] == it and np.abs(prob - want).max() <= TOL
    for k in (1, 2):  # single iterations: intermediate states agree as well
        w, _ = po.expectation_maximization(f, ident, 0.01, start, max_iterations=k)
        g = api.expectation_maximization(fdev, ident, 8, 0.01, start, ctx=gpu_ctx, max_iterations=k)
        assert np.abs(g - w).max() <= TOL
    fdev.free()


def test_many_cells(gpu_ctx):
    """14 000 cells, close to the 14-bit group id limit of PosData"""
    n = 14000
    cfg = SynthConfig(n_cells=n, coverage=0.02, n_loci=600, n_chr=1, n_clones=2, theta=0.01, seed=5)
    p = make_pileup(cfg)
    ident = np.arange(n, dtype=np.uint32)
    start = np.random.default_rng(2).uniform(0.2, 0.8, n)
    want, it = po.expectation_maximization(p, ident, 0.01, start, max_iterations=3)
    got, st = api.expectation_maximization(p, ident, 8, 0.01, start, ctx=gpu_ctx, max_iterations=3, return_stats=True)
    assert st["iterations"] == it and np.abs(got - want).max() <= TOL


def test_out_of_range_groups_are_an_error(gpu_ctx):
    p = repeated_locus([5 << 2 | 1, 1 << 2 | 1])
    with pytest.raises(api.SgpuError) as e:
        api.expectation_maximization(p, np.arange(6, dtype=np.uint32), 1, THETA, [0.5, 0.5], ctx=gpu_ctx)
    assert e.value.code == -3
    m = np.array([0, NO_POS], np.uint32)  # log_likelihood.at(NO_POS) throws in the reference
    with pytest.raises(api.SgpuError):
        api.expectation_maximization(repeated_locus([0 << 2 | 1, 1 << 2 | 1]), m, 1, THETA, [0.5, 0.5], ctx=gpu_ctx)


def test_full_size_throughput_and_properties(gpu_ctx):
    """8 000 cells x 16 384 loci (cfg3 shape, ~33 M entries after the filter): the probabilities stay in [0, 1],
    a second run is bit-identical, and swapping the roles of the clusters (p -> 1 - p) mirrors the result"""
    n = 8000
    dev = gpu_ctx.synth_pileup(n, 0.5, 1, 16384, n_clones=2, theta=0.001, p_multi=0.005, p_mate=0.01, seed=3)
    ident = np.arange(n, dtype=np.uint32)
    fdev, _ = api.Filter(0.001, 4, gpu_ctx).filter_device(dev, ident)
    rng = np.random.default_rng(0)
    start = rng.uniform(0.3, 0.7, n)
    a, st = api.expectation_maximization(fdev, ident, 8, 0.001, start, ctx=gpu_ctx, max_iterations=4, return_stats=True)
    b = api.expectation_maximization(fdev, ident, 8, 0.001, start, ctx=gpu_ctx, max_iterations=4)
    assert np.array_equal(a, b) and a.min() >= 0.0 and a.max() <= 1.0
    c = api.expectation_maximization(fdev, ident, 8, 0.001, 1.0 - start, ctx=gpu_ctx, max_iterations=4)
    assert np.abs((1.0 - c) - a).max() <= 1e-6
    print("EM at 8000 cells:", st, "entries", fdev.n_entries, "loci", fdev.n_loci)
    fdev.free()
