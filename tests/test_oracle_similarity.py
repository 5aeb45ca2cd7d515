"""The oracle's similarity-matrix half against golden matrices produced by the compiled reference
(the reference's own tests do not pin this function numerically) and, when oracle/_ref
is present, against the compiled reference on fresh random inputs."""
import numpy as np
import pytest

from conftest import assert_matrix_close, golden_pileup, load_golden
from oracle import pyoracle as po
from secedo_b200.pileup import Pileup
from secedo_b200.synth import SynthConfig, make_pileup

CASES = ["sim_ten_rows", "sim_six_cells", "sim_three_rows", "sim_reference_test_style", "sim_synth_multi",
         "sim_synth_subcluster"]


@pytest.mark.parametrize("name", CASES)
def test_golden_matrices(name):
    g = load_golden(name)
    p = golden_pileup(g)
    for t in g["threads"]:
        for norm in g["norms"]:
            r = po.similarity(p, int(g["num_cells"]), int(g["L"]), g["gmap"], float(g["eps"]), float(g["h"]),
                              float(g["theta"]), int(t), str(norm))
            assert_matrix_close(r.M, g[f"M_t{t}_{norm}"], 1e-9)


def test_six_cells_is_all_zero():
    # everything is "tail": the reference never compares any read
    g = load_golden("sim_six_cells")
    assert not g["M_t1_ADD_MIN"].any()
    r = po.similarity(golden_pileup(g), int(g["num_cells"]), int(g["L"]), g["gmap"], 0.01, 0.5, 0.01, 1)
    assert not r.S1.any() and not r.D1.any() and int(r.K[0]) == 0


def test_log_prob_tables_golden():
    g = load_golden("log_probs")
    i = 0
    while f"params{i}" in g.files:
        e, h, t, L = g[f"params{i}"]
        n = g[f"ls{i}"].shape[0]
        ls, ld = po.log_probs(e, h, t, int(L), n)
        assert np.array_equal(ls, g[f"ls{i}"], equal_nan=True) and np.array_equal(ld, g[f"ld{i}"], equal_nan=True)
        i += 1
    assert i >= 5


def test_log_prob_appendix_b():
    # SURVEY.md Appendix B (values printed by the compiled reference with %.17g)
    ls, ld = po.log_probs(0.01, 0.5, 0.01, 1000, 4)
    assert ls[1, 0] == -0.30220118877320645 and ld[1, 0] == -0.30549911676764557
    assert ls[0, 1] == -1.3439605499705118 and ld[0, 1] == -1.3346722319155464
    assert ls[2, 1] == -1.5349456234505463 and ld[2, 1] == -1.5044997536033944
    assert ls[0, 3] == -2.7620170814746268 and ld[0, 3] == -2.7619593694533164


def test_decomposition_identity():
    """M_raw == F10*S1 + F01*D1 + sum G(s,d)*N_sd — the identity the CUDA path relies on."""
    g = load_golden("sim_synth_multi")
    p = golden_pileup(g)
    n = int(g["num_cells"])
    r = po.similarity(p, n, 1000, g["gmap"], 0.01, 0.5, 0.01, 2)
    ls, ld = po.log_probs(0.01, 0.5, 0.01, 1000, 16)
    F = ld - ls
    assert r.class_hist[:, :].sum() > 0 and r.class_hist[2:, :].sum() + r.class_hist[:, 2:].sum() > 0
    # only classes of order 2 occur per cell pair in H; higher orders are checked through the histogram total
    hist_hi = r.class_hist.copy()
    hist_hi[:3, :3] = 0
    if hist_hi.sum() == 0:
        G = lambda s, d: F[s, d] - s * F[1, 0] - d * F[0, 1]
        M = F[1, 0] * r.S1 + F[0, 1] * r.D1 + G(2, 0) * r.H[0] + G(1, 1) * r.H[1] + G(0, 2) * r.H[2]
        assert_matrix_close(M, r.raw, 1e-12)
    # totals: every counted pair contributes x_s to S1 and x_d to D1
    s_tot = sum(int(r.class_hist[s, d]) * s for s in range(16) for d in range(16))
    d_tot = sum(int(r.class_hist[s, d]) * d for s in range(16) for d in range(16))
    assert int(np.triu(r.S1, 1).sum()) == s_tot and int(np.triu(r.D1, 1).sum()) == d_tot


def test_normalize_invalid():
    with pytest.raises(ValueError):
        po.oracle_lib()
        import ctypes as C
        m = np.zeros((2, 2))
        rc = po.oracle_lib().orc_normalize(C.c_int(7), C.c_uint32(2), m.ctypes.data_as(C.POINTER(C.c_double)))
        if rc:
            raise ValueError("invalid normalization")


def test_fragment_span_rejected():
    # the same read id 1500 bp apart with L = 1000: undefined (batch-timing dependent) in the reference
    p = Pileup.from_pos_data([[(100, [1, 2], [0 << 2, 1 << 2]), (1600, [1, 3], [0 << 2, 2 << 2])]])
    with pytest.raises(ValueError):
        po.similarity(p, 3, 1000, np.arange(3), 0.01, 0.5, 0.01, 1)


@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("threads", [1, 2, 5])
def test_against_compiled_reference(threads):
    cfg = SynthConfig(n_cells=30, coverage=0.4, n_loci=400, n_chr=2, p_multi=0.5, p_mate=0.2, theta=0.03,
                      seed=100 + threads)
    p = make_pileup(cfg)
    ident = np.arange(cfg.n_cells, dtype=np.uint32)
    kl, ke, _, _ = po.filter_flags(p, ident, 0.01)
    f = p.select(kl, ke)
    for norm in ("ADD_MIN", "EXPONENTIATE", "SCALE_MAX_1"):
        M, _ = po.ref_similarity(f, cfg.n_cells, 1000, ident, 0.01, 0.5, 0.01, threads, norm)
        r = po.similarity(f, cfg.n_cells, 1000, ident, 0.01, 0.5, 0.01, threads, norm)
        assert_matrix_close(r.M, M, 1e-9)


def test_wide_restatement_equals_the_pinned_one():
    """BASELINE config 5 (20 000 cells) is beyond the reference's 14-bit group ids, so its oracle is the restatement on
    32-bit entries. Below 16 384 groups the two entry widths must agree exactly (same code, other loads), which ties the
    wide oracle to the compiled reference through the 16-bit one; and group ids >= 16 384 are carried through."""
    from secedo_b200.pileup import NO_POS, NO_POS_WIDE, Pileup
    cfg = SynthConfig(n_cells=50, coverage=0.4, n_loci=500, n_chr=2, p_multi=0.3, p_mate=0.1, theta=0.02, seed=41)
    p = make_pileup(cfg)
    wide = Pileup(p.chr_ptr, p.row_ptr, p.position, p.read_id, p.gid_base.astype(np.uint32))
    assert wide.wide and not p.wide
    members = np.r_[0:20, 25:50]
    for no_pos, pile in ((NO_POS, p), (NO_POS_WIDE, wide)):
        gmap = np.full(50, no_pos, np.uint32)
        gmap[members] = np.arange(members.size)
        kl, ke, cov, cov64 = po.filter_flags(pile, gmap, 0.01)
        f = pile.select(kl, ke)
        o = po.similarity(f, members.size, 1000, gmap, 0.01, 0.5, 0.01, 3, "ADD_MIN")
        if no_pos == NO_POS:
            ref = (kl, ke, cov64, o)
        else:
            assert np.array_equal(kl, ref[0]) and np.array_equal(ke, ref[1]) and cov64 == ref[2]
            for name in ("M", "S1", "D1", "H", "class_hist", "K", "raw"):
                assert np.array_equal(getattr(o, name), getattr(ref[3], name)), name
    # the same reads under group ids shifted beyond 14 bits give the same matrix
    shift = 20000
    shifted = Pileup(p.chr_ptr, p.row_ptr, p.position, p.read_id,
                     ((p.gid_base.astype(np.uint32) >> 2) + shift) << 2 | (p.gid_base.astype(np.uint32) & 3))
    gmap = np.full(shift + 50, NO_POS_WIDE, np.uint32)
    gmap[shift + members] = np.arange(members.size)
    kl, ke, _, _ = po.filter_flags(shifted, gmap, 0.01)
    assert np.array_equal(kl, ref[0]) and np.array_equal(ke, ref[1])
    o = po.similarity(shifted.select(kl, ke), members.size, 1000, gmap, 0.01, 0.5, 0.01, 3, "ADD_MIN")
    assert np.array_equal(o.M, ref[3].M) and np.array_equal(o.S1, ref[3].S1)


def test_span_chain_golden():
    """a read id chained over >= L bp: the compiled reference (golden) equals the restatement on the pileup whose chain
    was relabelled into the reads the reference makes of it (retire at start + L <= position, new read at the next entry)
    — the semantics the CUDA path implements by splitting; the restatement itself refuses the un-split input"""
    g = load_golden("sim_span_chain")
    p = golden_pileup(g)
    split = Pileup(p.chr_ptr, p.row_ptr, p.position, g["split_read_id"], p.gid_base)
    assert (p.read_id != split.read_id).sum() == 9
    for t in g["threads"]:
        o = po.similarity(split, int(g["num_cells"]), int(g["L"]), g["gmap"], float(g["eps"]), float(g["h"]), float(g["theta"]), int(t))
        assert_matrix_close(o.M, g[f"M_t{t}_ADD_MIN"], 1e-9)
        with pytest.raises(ValueError):
            po.similarity(p, int(g["num_cells"]), int(g["L"]), g["gmap"], float(g["eps"]), float(g["h"]), float(g["theta"]), int(t))
