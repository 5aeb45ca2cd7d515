"""Wide pileups: group ids beyond the reference's 14 bits (BASELINE config 5: 20 000 cells). The oracle is the
restatement on 32-bit entries (tests/test_oracle_similarity.py ties it to the 16-bit one, which is pinned against the
compiled reference)."""
import numpy as np
import pytest

from conftest import assert_matrix_close
from oracle import pyoracle as po
from secedo_b200 import api
from secedo_b200.pileup import NO_POS_WIDE, Pileup
from secedo_b200.synth import SynthConfig, make_pileup

pytestmark = pytest.mark.gpu
L = 1000


def check(ctx, f, n_cells, gmap, eps, h, theta, threads, path):
    o = po.similarity(f, n_cells, L, gmap, eps, h, theta, threads, "ADD_MIN")
    c = api.Counts(ctx, n_cells)
    c.accumulate(f, L, gmap, eps, h, theta, threads, path)
    S1, D1, H, hist = c.download()
    assert np.array_equal(S1, o.S1), "S1 differs"
    assert np.array_equal(D1, o.D1), "D1 differs"
    assert np.array_equal(H, o.H), "H differs"
    oh = o.class_hist.copy()
    oh[0, 0] = oh[0, 1] = oh[1, 0] = 0
    assert np.array_equal(hist, oh)
    assert_matrix_close(c.finalize(L, eps, h, theta, "ADD_MIN"), o.M, 1e-6)
    c.free()
    return o


@pytest.mark.parametrize("path", ["scatter", "gemm"])
def test_wide_entries_small(gpu_ctx, path):
    """the same reads as 16-bit and as 32-bit entries, and under group ids shifted beyond 14 bits with a sub-cluster map"""
    cfg = SynthConfig(n_cells=300, coverage=0.3, n_loci=1500, n_chr=2, p_multi=0.3, p_mate=0.1, theta=0.02, seed=43)
    p = make_pileup(cfg)
    shift = 17000
    gb = p.gid_base.astype(np.uint32)
    shifted = Pileup(p.chr_ptr, p.row_ptr, p.position, p.read_id, (((gb >> 2) + shift) << 2) | (gb & 3))
    members = np.r_[0:120, 150:300]
    gmap = np.full(shift + 300, NO_POS_WIDE, np.uint32)
    gmap[shift + members] = np.arange(members.size)
    flt = api.Filter(0.01, 4, gpu_ctx)
    f, cov = flt.filter(shifted, gmap, "", 1)
    assert f.wide
    kl, ke, _, cov64 = po.filter_flags(shifted, gmap, 0.01)
    assert f == shifted.select(kl, ke) and cov == cov64
    o = check(gpu_ctx, f, members.size, gmap, 0.01, 0.5, 0.01, 4, path)
    # 16-bit twin: identical matrix
    gmap16 = np.full(300, 16383, np.uint32)
    gmap16[members] = np.arange(members.size)
    f16, _ = flt.filter(p, gmap16, "", 1)
    M16 = api.compute_similarity_matrix(f16, members.size, L, gmap16, 0.01, 0.5, 0.01, 4, "", "ADD_MIN", ctx=gpu_ctx, path=path)
    assert_matrix_close(M16, o.M, 1e-6)


def test_wide_pieces_and_cutoff(gpu_ctx):
    """pieces of chromosomes and the cutoff from the chromosome ends on wide entries"""
    from secedo_b200.dist import plan_pieces
    from test_gpu_pieces import chrom_positions, cutoffs_from_ends, piece_pileup
    cfg = SynthConfig(n_cells=80, coverage=0.4, n_loci=600, n_chr=2, p_multi=0.4, p_mate=0.1, theta=0.02, seed=44)
    p = make_pileup(cfg)
    p = Pileup(p.chr_ptr, p.row_ptr, p.position, p.read_id, p.gid_base.astype(np.uint32))
    ident = np.arange(80, dtype=np.uint32)
    f, _ = api.Filter(0.01, 4, gpu_ctx).filter(p, ident, "", 1)
    o = po.similarity(f, 80, L, ident, 0.01, 0.5, 0.01, 2, "ADD_MIN")
    tail, _ = cutoffs_from_ends(gpu_ctx, f, 2)
    c = api.Counts(gpu_ctx, 80)
    for piece in plan_pieces(chrom_positions(f), 3, L):
        c.accumulate_range(piece_pileup(f, piece), L, ident, 0.01, 0.5, 0.01, [d["own_pos_begin"] for d in piece],
                           [d["own_pos_end"] for d in piece], [tail[d["chrom"]] for d in piece], "gemm")
    S1, D1, H, _ = c.download()
    assert np.array_equal(S1, o.S1) and np.array_equal(D1, o.D1) and np.array_equal(H, o.H)
    c.free()


def test_cfg5_20000_cells_vs_wide_oracle(gpu_ctx):
    """BASELINE.json configs[4] at its real cell count and coverage: 20 000 cells at 1x (20 000 reads, 2e8 cross-cell pairs
    per locus), count planes of 1.6 GB each, on a handful of loci the enumerating restatement finishes in about a minute.
    Integer counts bit-exact on both first-order paths, matrix within 1e-6 * max|M|."""
    n = 20000
    cfg = SynthConfig(n_cells=n, coverage=1.0, n_loci=4, n_chr=1, n_clones=2, p_multi=0.16, p_mate=0.02, theta=0.001, seed=9)
    p = make_pileup(cfg)
    assert p.wide
    ident = np.arange(n, dtype=np.uint32)
    f, _ = api.Filter(0.001, 4, gpu_ctx).filter(p, ident, "", 1)
    kl, ke, _, _ = po.filter_flags(p, ident, 0.001)
    assert f == p.select(kl, ke) and f.n_loci >= 2
    o = po.similarity(f, n, L, ident, 0.01, 0.15, 0.001, 8, "ADD_MIN")
    for path in ("gemm", "scatter"):
        c = api.Counts(gpu_ctx, n)
        c.accumulate(f, L, ident, 0.01, 0.15, 0.001, 8, path)
        S1, D1, H, hist = c.download()
        assert np.array_equal(S1, o.S1), f"{path}: S1"
        assert np.array_equal(D1, o.D1), f"{path}: D1"
        assert np.array_equal(H, o.H), f"{path}: H"
        del S1, D1, H
        if path == "gemm":
            assert_matrix_close(c.finalize(L, 0.01, 0.15, 0.001, "ADD_MIN"), o.M, 1e-6)
        c.free()
    assert o.S1.sum() > 10 ** 8


def test_cfg5_device_generated_paths_agree(gpu_ctx):
    """20 000 cells on the device generator (wide entries), more loci than the oracle can follow: the tcgen05 path and
    the pair scatter agree bit for bit, two chromosomes one after the other equal both at once"""
    n = 20000
    dev = gpu_ctx.synth_pileup(n, 0.25, 2, 64, n_clones=4, theta=0.001, p_multi=0.02, p_mate=0.01, seed=19)
    assert dev.wide
    ident = np.arange(n, dtype=np.uint32)
    fdev, _ = api.Filter(0.001, 4, gpu_ctx).filter_device(dev, ident)
    assert fdev.wide and fdev.n_loci > 20
    sums = {}
    for path in ("gemm", "scatter"):
        c = api.Counts(gpu_ctx, n)
        st = c.accumulate(fdev, L, ident, 0.01, 0.5, 0.001, 8, path)
        assert st["path_used"] == path
        sums[path] = c.checksum()
        if path == "gemm":
            M = c.finalize(L, 0.01, 0.5, 0.001, "ADD_MIN")
            assert np.array_equal(M, M.T) and not np.diag(M).any() and M.min() == 0.0 and np.isfinite(M).all()
            del M
        c.free()
    assert sums["gemm"] == sums["scatter"] != 0
    f = fdev.download()
    two = api.Counts(gpu_ctx, n)
    for c_ in range(2):
        two.accumulate(f.loci_range(c_, 0, 1 << 40), L, ident, 0.01, 0.5, 0.001, 8, "gemm")
    assert two.checksum() == sums["gemm"]
    two.free()
