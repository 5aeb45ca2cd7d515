"""K1 parity: the CUDA filter (through the C ABI) against the oracle and the golden vectors."""
import numpy as np
import pytest

from conftest import golden_pileup, load_golden
from oracle import pyoracle as po
from secedo_b200 import api
from secedo_b200.api import Filter
from secedo_b200.pileup import NO_POS, Pileup
from secedo_b200.synth import SynthConfig, make_pileup

pytestmark = pytest.mark.gpu


def test_is_significant_golden_tuples(gpu_ctx):
    g = load_golden("filter_tuples")
    for key in g.files:
        if not key.startswith("sig_"):
            continue
        theta = float(key.split("theta")[1].split("_")[0])
        cp = int(key.split("cp")[1])
        got = Filter(theta, cp, gpu_ctx).is_significant(g["tuples"])
        assert np.array_equal(got.astype(np.uint8), g[key]), key


def test_is_significant_dense_sweep(gpu_ctx):
    """every count tuple (sorted) with coverage <= 60, plus random high-coverage ones: bit-exact
    boolean decisions vs the oracle (pow/log differ by ulps between glibc and CUDA)."""
    t = [(a, b, c, d) for d in range(0, 61) for c in range(0, min(d, 20) + 1) for b in range(0, min(c, 6) + 1)
         for a in range(0, min(b, 3) + 1)]
    t = np.array(t, np.uint16)
    rng = np.random.default_rng(1)
    hi = np.stack([rng.integers(0, 8, 20000), rng.integers(0, 30, 20000), rng.integers(0, 3000, 20000),
                   rng.integers(100, 20000, 20000)], 1).astype(np.uint16)
    t = np.concatenate([t, hi])
    for theta in (0.01, 0.001, 0.05):
        got = Filter(theta, 4, gpu_ctx).is_significant(t)
        assert np.array_equal(got.astype(np.uint8), po.is_significant(t, theta, 4)), theta


def test_reference_known_answers(gpu_ctx):
    # tests/test_is_significant.cpp:46-84
    def sig(bases, theta=0.01):
        c = [bases.upper().count(x) for x in "ACGT"]
        return Filter(theta, 4, gpu_ctx).is_significant(c)
    assert not sig("C" * 43 + "T" + "C" * 8)
    assert sig("CCACGTACGTACCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCTCCCCCCCC")
    assert not sig("tttttTTTTaTTTttTaTtTTTTtTTtTTTtTttTTtTtTtttTTttttTTttTTTTtt")
    assert not sig("CcccccccCcccCCCCcaCcccCccACccccCCCcCcCCCC", 0.001)
    assert not sig("GGG" + "A" * 118)
    assert not sig("GGGA", 0.001)


def test_filter_reference_cases(gpu_ctx):
    # tests/test_is_significant.cpp:96-216
    def assemble(pos, read_ids, cell_ids, bases):
        return (pos, list(read_ids), [(c << 2) | b for c, b in zip(cell_ids, bases)])
    pd = assemble(1, [2 * i + 1 for i in range(100)], range(100), [0 if i < 10 else 1 for i in range(100)])
    flt = Filter(1e-3, 4, gpu_ctx)
    f, cov = flt.filter(Pileup.empty(0), np.zeros(0, np.uint32), "", 1)
    assert f.n_chr == 0 and f.n_loci == 0
    p = Pileup.from_pos_data([[pd]])
    f, cov = flt.filter(p, np.arange(100), "", 1)
    assert cov == 100.0 and f == p
    p = Pileup.from_pos_data([[(i + 1, pd[1], pd[2]) for i in range(100)] for _ in range(23)])
    f, cov = flt.filter(p, np.arange(100), "", 2)
    assert cov == 100.0 and f.n_chr == 23 and f == p
    nd = assemble(1, [0, 5, 9], [1, 3, 5], [0, 0, 0])
    p = Pileup.from_pos_data([[(i + 1, nd[1], nd[2]) for i in range(100)]])
    f, cov = flt.filter(p, np.arange(10), "", 2)
    assert cov == 0 and f.n_chr == 1 and f.n_loci == 0


@pytest.mark.parametrize("prefix,map_key,cov_key", [("f_", "id_to_pos", "avg_coverage"),
                                                    ("sub_f_", "sub_id_to_pos", "sub_avg_coverage")])
def test_filter_golden_passthrough(gpu_ctx, prefix, map_key, cov_key):
    g = load_golden("filter_synth")
    f, cov = Filter(float(g["theta"]), 4, gpu_ctx).filter(golden_pileup(g), g[map_key], "", 1)
    assert f == golden_pileup(g, prefix)
    assert cov == float(g[cov_key])


@pytest.mark.parametrize("n_cells,coverage,n_loci,n_chr", [(50, 0.05, 3000, 2), (500, 0.05, 4000, 1),
                                                           (2000, 0.1, 1500, 3), (300, 2.0, 800, 24)])
def test_filter_vs_oracle(gpu_ctx, n_cells, coverage, n_loci, n_chr):
    cfg = SynthConfig(n_cells=n_cells, coverage=coverage, n_loci=n_loci, n_chr=n_chr, seed=n_cells)
    p = make_pileup(cfg)
    rng = np.random.default_rng(2)
    for sub in (False, True):
        id_to_pos = np.arange(n_cells, dtype=np.uint32)
        if sub:
            id_to_pos[rng.random(n_cells) < 0.5] = NO_POS
        for theta in (0.01, 0.001):
            kl, ke, cov, cov64 = po.filter_flags(p, id_to_pos, theta)
            f, gcov = Filter(theta, 4, gpu_ctx).filter(p, id_to_pos, "", 1)
            assert f == p.select(kl, ke)
            assert gcov == cov64


def test_filter_rejects_bad_group(gpu_ctx):
    from secedo_b200.api import SgpuError
    p = Pileup.from_pos_data([[(1, [1, 2], [(7 << 2) | 1, (2 << 2) | 1])]])
    with pytest.raises(SgpuError):
        Filter(0.01, 4, gpu_ctx).filter(p, np.arange(4), "", 1)


def test_lazy_upload_pulls_read_ids_of_kept_loci(gpu_ctx):
    """sgpu_pileup_upload_lazy_async: the read ids stay in pinned host memory and the filter's compaction reads those
    of the kept loci through the mapped pointer - same filtered pileup bit for bit; other consumers copy them first"""
    import torch
    cfg = SynthConfig(n_cells=300, coverage=0.3, n_loci=2500, n_chr=3, p_multi=0.2, p_mate=0.05, theta=0.01, seed=12)
    p = make_pileup(cfg)
    pinned = {k: torch.from_numpy(np.ascontiguousarray(getattr(p, k))).pin_memory() for k in ("row_ptr", "position", "read_id", "gid_base")}
    hp = Pileup(p.chr_ptr, *(pinned[k].numpy() for k in ("row_ptr", "position", "read_id", "gid_base")))
    ident = np.arange(cfg.n_cells, dtype=np.uint32)
    flt = api.Filter(0.01, 4, gpu_ctx)
    want, cov = flt.filter(p, ident, "", 1)
    lazy = gpu_ctx.upload_lazy_async(hp)
    fdev, cov2 = flt.filter_device(lazy, ident)
    assert fdev.download() == want and cov2 == cov
    # sub-cluster: entries of other cells are dropped while compacting
    sub = np.full(cfg.n_cells, NO_POS, np.uint32)
    sub[40:200] = np.arange(160)
    want_sub, _ = flt.filter(p, sub, "", 1)
    fsub, _ = flt.filter_device(lazy, sub)
    assert fsub.download() == want_sub
    # any other consumer materialises the read ids: similarity counts straight from the lazy pileup, and its download
    c1, c2 = api.Counts(gpu_ctx, cfg.n_cells), api.Counts(gpu_ctx, cfg.n_cells)
    lazy2 = gpu_ctx.upload_lazy_async(hp)
    c1.accumulate(lazy2, 1000, ident, 0.01, 0.5, 0.01, 8, "gemm")
    c2.accumulate(p, 1000, ident, 0.01, 0.5, 0.01, 8, "gemm")
    for a, b in zip(c1.download(), c2.download()):
        assert np.array_equal(a, b)
    assert lazy2.download() == p and gpu_ctx.upload_lazy_async(hp).download() == p
    with pytest.raises(api.SgpuError):
        gpu_ctx.upload_lazy_async(p)  # pageable memory cannot be read by the device
    for o in (fdev, fsub, lazy, lazy2, c1, c2):
        o.free()
