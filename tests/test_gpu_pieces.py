"""Chromosomes accumulated in pieces (locus ranges with a halo): the sums over the pieces equal the whole
— integer counts bit for bit against the oracle — and the tail cutoff decided from the END of a chromosome
(sgpu_chromosome_cutoff) equals the cutoff of the whole chromosome."""
import numpy as np
import pytest

from conftest import assert_matrix_close
from oracle import pyoracle as po
from secedo_b200 import api
from secedo_b200.dist import NO_TAIL, plan_pieces
from secedo_b200.pileup import Pileup
from secedo_b200.synth import SynthConfig, make_pileup

pytestmark = pytest.mark.gpu
L = 1000


def chrom_positions(f):
    return [f.position[int(f.chr_ptr[c]):int(f.chr_ptr[c + 1])] for c in range(f.n_chr)]


def piece_pileup(f, piece):
    """host pileup of one piece: per touched chromosome the loci [lo, hi) (owned + halo)"""
    return Pileup.concat([f.loci_range(d["chrom"], d["lo"], d["hi"]) for d in piece])


def cutoffs_from_ends(ctx, f, threads, first_suffix_bp=6000):
    """tail position of every chromosome from its end, with longer and longer suffixes until resolved"""
    pos = chrom_positions(f)
    tail = np.full(f.n_chr, NO_TAIL, np.uint32)
    tries = 0
    for c in range(f.n_chr):
        n = len(pos[c])
        if n == 0:
            continue
        span = first_suffix_bp
        while True:
            lo = int(np.searchsorted(pos[c], int(pos[c][-1]) - span, side="left"))
            whole = lo == 0
            tp, ok = api.chromosome_cutoff(f.loci_range(c, lo, n), L, threads, [whole], ctx=ctx)
            tries += 1
            if ok[0]:
                tail[c] = tp[0]
                break
            assert not whole, "a whole chromosome always resolves"
            span *= 2
    return tail, tries


@pytest.mark.parametrize("path", ["scatter", "gemm"])
@pytest.mark.parametrize("threads,n_pieces", [(1, 3), (2, 2), (8, 5)])
def test_pieces_equal_whole(gpu_ctx, threads, n_pieces, path):
    cfg = SynthConfig(n_cells=70, coverage=0.4, n_loci=700, n_chr=3, p_multi=0.45, p_mate=0.15, theta=0.02, seed=11)
    p = make_pileup(cfg)
    # uneven chromosomes, one of them tiny
    p = Pileup.concat([p.loci_range(0, 0, 700), p.loci_range(1, 0, 40), p.loci_range(2, 0, 333)])
    ident = np.arange(cfg.n_cells, dtype=np.uint32)
    f, _ = api.Filter(0.01, 4, gpu_ctx).filter(p, ident, "", 1)
    o = po.similarity(f, cfg.n_cells, L, ident, 0.01, 0.5, 0.01, threads, "ADD_MIN")
    # the cutoff from the ends == the cutoff of the whole chromosomes
    tail, _ = cutoffs_from_ends(gpu_ctx, f, threads)
    tp_whole, ok = api.chromosome_cutoff(f, L, threads, [True] * f.n_chr, ctx=gpu_ctx)
    assert ok.all() and np.array_equal(tail, tp_whole)
    pieces = plan_pieces(chrom_positions(f), n_pieces, L)
    c = api.Counts(gpu_ctx, cfg.n_cells)
    multi = 0
    for piece in pieces:
        if not piece:
            continue
        st = c.accumulate_range(piece_pileup(f, piece), L, ident, 0.01, 0.5, 0.01,
                                [d["own_pos_begin"] for d in piece], [d["own_pos_end"] for d in piece],
                                [tail[d["chrom"]] for d in piece], path)
        multi += st["n_pairs_multi"]
    S1, D1, H, hist = c.download()
    assert np.array_equal(S1, o.S1), "S1 differs"
    assert np.array_equal(D1, o.D1), "D1 differs"
    assert np.array_equal(H, o.H), "second-order class counts differ"
    oh = o.class_hist.copy()
    oh[0, 0] = oh[0, 1] = oh[1, 0] = 0
    assert np.array_equal(hist, oh) and multi == int(oh.sum())
    assert_matrix_close(c.finalize(L, 0.01, 0.5, 0.01, "ADD_MIN"), o.M, 1e-6)
    assert o.H.sum() > 0 and (o.K < np.diff(f.chr_ptr.astype(np.int64)) * 1000).all()
    c.free()


def test_tail_auto_for_the_piece_that_holds_the_end(gpu_ctx):
    """TAIL_AUTO: the piece that holds the end of a chromosome decides the cutoff itself; the other pieces are told that
    none of their reads are tail reads of ... the end piece's range (they get the position from the same decision made
    by sgpu_chromosome_cutoff). A piece too short for the decision is refused."""
    cfg = SynthConfig(n_cells=70, coverage=0.4, n_loci=600, n_chr=2, p_multi=0.4, p_mate=0.1, theta=0.02, seed=12)
    ident = np.arange(cfg.n_cells, dtype=np.uint32)
    f, _ = api.Filter(0.01, 4, gpu_ctx).filter(make_pileup(cfg), ident, "", 1)
    T = 2
    o = po.similarity(f, cfg.n_cells, L, ident, 0.01, 0.5, 0.01, T, "ADD_MIN")
    tail, _ = cutoffs_from_ends(gpu_ctx, f, T)
    # every chromosome cut in its middle: piece 0 = the first halves, piece 1 = the second halves (with their halos)
    pieces = [[], []]
    for c_, pos in enumerate(chrom_positions(f)):
        one = plan_pieces([pos], 2, L)
        for k in range(2):
            d = dict(one[k][0])
            d["chrom"] = c_
            pieces[k].append(d)
    c = api.Counts(gpu_ctx, cfg.n_cells)
    for piece in pieces:
        tp = [api.TAIL_AUTO if d["own_pos_end"] == NO_TAIL else tail[d["chrom"]] for d in piece]
        c.accumulate_range(piece_pileup(f, piece), L, ident, 0.01, 0.5, 0.01, [d["own_pos_begin"] for d in piece],
                           [d["own_pos_end"] for d in piece], tp, "gemm", num_threads=T)
    S1, D1, H, _ = c.download()
    assert np.array_equal(S1, o.S1) and np.array_equal(D1, o.D1) and np.array_equal(H, o.H)
    c.free()
    # the last two loci of a chromosome cannot decide its cutoff
    n0 = int(f.chr_ptr[1])
    tiny = f.loci_range(0, n0 - 2, n0)
    c = api.Counts(gpu_ctx, cfg.n_cells)
    with pytest.raises(api.SgpuError):
        c.accumulate_range(tiny, L, ident, 0.01, 0.5, 0.01, [int(tiny.position[0])], [NO_TAIL], [api.TAIL_AUTO], "scatter", num_threads=T)
    c.free()


def test_pieces_second_order_gemm(gpu_ctx, monkeypatch):
    """pieces with the second-order GEMM forced: the locus pairs belong to the piece that owns their first locus, the
    pairs of order >= 3 are taken back by the piece that owns their first common locus; the sums are exact"""
    monkeypatch.setenv("SECEDO_B200_SECOND_ORDER", "gemm")
    cfg = SynthConfig(n_cells=60, coverage=0.3, n_loci=900, n_chr=2, p_multi=0.45, p_mate=0.15, theta=0.02, seed=7)
    ident = np.arange(60, dtype=np.uint32)
    f, _ = api.Filter(0.01, 4, gpu_ctx).filter(make_pileup(cfg), ident, "", 1)
    o = po.similarity(f, 60, L, ident, 0.01, 0.5, 0.01, 2, "ADD_MIN")
    tail, _ = cutoffs_from_ends(gpu_ctx, f, 2)
    c = api.Counts(gpu_ctx, 60)
    for piece in plan_pieces(chrom_positions(f), 4, L):
        c.accumulate_range(piece_pileup(f, piece), L, ident, 0.01, 0.5, 0.01, [d["own_pos_begin"] for d in piece],
                           [d["own_pos_end"] for d in piece], [tail[d["chrom"]] for d in piece], "gemm")
    S1, D1, H, hist = c.download()
    assert np.array_equal(S1, o.S1) and np.array_equal(D1, o.D1) and np.array_equal(H, o.H)
    assert_matrix_close(c.finalize(L, 0.01, 0.5, 0.01, "ADD_MIN"), o.M, 1e-6)
    c.free()


def test_cutoff_suffix_too_short_is_reported(gpu_ctx):
    """a batch size (4 * num_threads) that no stretch of the suffix can fill: no batch trigger fires 'for sure' there, so
    the cutoff depends on the history before the suffix; the suffix must say so, the whole chromosome resolves"""
    cfg = SynthConfig(n_cells=30, coverage=0.2, n_loci=400, n_chr=1, p_multi=0.2, seed=3)
    p = make_pileup(cfg)
    T = 300  # 1 200 completed reads per batch: about 200 loci's worth
    tp, ok = api.chromosome_cutoff(p.loci_range(0, 320, p.n_loci), L, T, [False], ctx=gpu_ctx)
    assert not ok[0]
    tp, ok = api.chromosome_cutoff(p, L, T, [True], ctx=gpu_ctx)
    assert ok[0]
    ident = np.arange(30, dtype=np.uint32)
    o = po.similarity(p, 30, L, ident, 0.01, 0.5, 0.01, T, "ADD_MIN")
    assert 0 < o.K[0] < p.n_entries
    # one piece = the whole chromosome, cutoff passed in: equals the plain call and the oracle
    c = api.Counts(gpu_ctx, 30)
    c.accumulate_range(p, L, ident, 0.01, 0.5, 0.01, [0], [NO_TAIL], tp, "scatter")
    S1, D1, H, _ = c.download()
    assert np.array_equal(S1, o.S1) and np.array_equal(D1, o.D1) and np.array_equal(H, o.H)
    c.free()


def test_full_size_pieces(gpu_ctx):
    """8 000 cells at 0.5x on the device generator: two pieces of each chromosome == the whole, on the tcgen05 path"""
    dev = gpu_ctx.synth_pileup(8000, 0.5, 2, 256, n_clones=2, theta=0.001, p_multi=0.05, p_mate=0.01, seed=31)
    ident = np.arange(8000, dtype=np.uint32)
    fdev, _ = api.Filter(0.001, 4, gpu_ctx).filter_device(dev, ident)
    whole = api.Counts(gpu_ctx, 8000)
    whole.accumulate(fdev, L, ident, 0.01, 0.15, 0.001, 8, "gemm")
    f = fdev.download()
    tail, tries = cutoffs_from_ends(gpu_ctx, f, 8)
    assert tries == f.n_chr, "dense pileup: the first short suffix decides"
    parts = api.Counts(gpu_ctx, 8000)
    for piece in plan_pieces(chrom_positions(f), 2, L):
        parts.accumulate_range(piece_pileup(f, piece), L, ident, 0.01, 0.15, 0.001, [d["own_pos_begin"] for d in piece],
                               [d["own_pos_end"] for d in piece], [tail[d["chrom"]] for d in piece], "gemm")
    for a, b, name in zip(whole.download(), parts.download(), ("S", "D", "H", "hist")):
        assert np.array_equal(a, b), name
    whole.free()
    parts.free()
