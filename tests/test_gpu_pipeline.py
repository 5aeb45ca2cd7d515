"""The pieces chained the way divide_cluster chains them (spectral_clustering.cpp:336-377), every stage on the GPU and
nothing but the eigenvectors / probabilities leaving it: Filter::filter -> computeSimilarityMatrix -> laplacian +
eig_sym -> Fiedler split (ClusteringType::FIEDLER, :212-224) -> expectation_maximization. The planted two-clone
structure of the synthetic pileup must come back, and every stage is compared with the oracle on the way."""
import numpy as np
import pytest

from oracle import pyoracle as po
from secedo_b200 import api
from secedo_b200.synth import SynthConfig, clone_assignment, make_pileup

pytestmark = pytest.mark.gpu


def test_two_clones_recovered_end_to_end(gpu_ctx):
    cfg = SynthConfig(n_cells=400, coverage=0.5, n_loci=6000, n_chr=3, n_clones=2, frac_somatic=0.3, frac_germline=0.1,
                      p_multi=0.05, p_mate=0.02, theta=0.01, seed=77)
    truth = clone_assignment(cfg)
    n = cfg.n_cells
    ident = np.arange(n, dtype=np.uint32)
    theta, eps, h, L, threads = 0.01, 0.01, 0.5, 1000, 8
    dev = gpu_ctx.upload(make_pileup(cfg))
    fdev, coverage = api.Filter(theta, 4, gpu_ctx).filter_device(dev, ident)
    assert fdev.n_loci > 200
    counts = api.Counts(gpu_ctx, n)
    counts.accumulate(fdev, L, ident, eps, h, theta, threads, "auto")
    ev, vec, st, sim = counts.finalize_spectral(L, eps, h, theta, "ADD_MIN", k=7, want_matrix=True)
    # oracle on the same filtered pileup
    f = fdev.download()
    o = po.similarity(f, n, L, ident, eps, h, theta, threads, "ADD_MIN")
    assert np.abs(sim - o.M).max() <= 1e-6 * np.abs(o.M).max()
    w, v = po.spectral_embedding(o.M, 3)
    assert np.abs(ev[:3] - w).max() <= 1e-8
    # a real split: the second eigenvalue stands clear of the bulk
    assert ev[1] < 0.9 * ev[2]
    fiedler = vec[:, 1]
    assert min(np.linalg.norm(fiedler - v[:, 1]), np.linalg.norm(fiedler + v[:, 1])) <= 1e-6
    cluster = (fiedler >= 0).astype(np.float64)  # spectral_clustering.cpp:221-223
    agree = max(np.mean(cluster == truth), np.mean(cluster == 1 - truth))
    assert agree >= 0.97, agree
    # EM refinement of that split on the device-resident filtered pileup
    refined, est = api.expectation_maximization(fdev, ident, threads, theta, cluster, ctx=gpu_ctx, return_stats=True)
    want, it = po.expectation_maximization(f, ident, theta, cluster)
    assert est["iterations"] == it and np.abs(refined - want).max() <= 1e-6
    labels = refined > 0.5
    agree_em = max(np.mean(labels == truth), np.mean(labels == 1 - truth))
    assert agree_em >= agree and agree_em >= 0.99, (agree, agree_em)
    for obj in (counts, fdev, dev):
        obj.free()
