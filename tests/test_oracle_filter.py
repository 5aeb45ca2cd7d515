"""The oracle's filter half against (1) the reference's own known-answer tests
(tests/test_is_significant.cpp:46-216, replayed here with the same inputs), (2) golden decisions and
pass-through produced by the compiled reference (tests/golden/make_golden.py), (3) the compiled
reference itself when oracle/_ref is present."""
import numpy as np
import pytest

from conftest import golden_pileup, load_golden
from oracle import pyoracle as po
from secedo_b200.pileup import Pileup

CHAR = {"A": 0, "C": 1, "G": 2, "T": 3}


def sig(bases, theta=0.01):
    c = np.zeros(4, np.uint16)
    for ch in bases.upper():
        c[CHAR[ch]] += 1
    return bool(po.is_significant(c, theta)[0])


# ---- tests/test_is_significant.cpp:46-84 -------------------------------------------------------------
def test_cov52_one_different():
    assert not sig("CCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCTCCCCCCCC")


def test_cov52_ten_different():
    assert sig("CCACGTACGTACCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCTCCCCCCCC")


def test_cov59_two_different():
    assert not sig("tttttTTTTaTTTttTaTtTTTTtTTtTTTtTttTTtTtTtttTTttttTTttTTTTtt")


def test_at_limit():
    assert not sig("CcccccccCcccCCCCcaCcccCccACccccCCCcCcCCCC", 0.001)


def test_two_sigmas_away_false():
    assert not sig("GGG" + "A" * 118, 0.01)


def test_paradox():
    assert not sig("GGGA", 0.001)


# ---- tests/test_is_significant.cpp:96-216 (Filter::filter) ------------------------------------------
def assemble(pos, read_ids, cell_ids, bases):
    return (pos, list(read_ids), [(c << 2) | b for c, b in zip(cell_ids, bases)])


def run_filter(p, id_to_pos, theta):
    kl, ke, cov, cov64 = po.filter_flags(p, id_to_pos, theta)
    return p.select(kl, ke), cov


def test_filter_empty():
    f, cov = run_filter(Pileup.empty(0), np.zeros(0, np.uint32), 1e-3)
    assert f.n_chr == 0 and f.n_loci == 0


def one_pos(coverage=100, num_diff=10):
    return assemble(1, [2 * i + 1 for i in range(coverage)], range(coverage),
                    [0 if i < num_diff else 1 for i in range(coverage)])


def test_filter_one_pos_significant():
    p = Pileup.from_pos_data([[one_pos()]])
    f, cov = run_filter(p, np.arange(100), 1e-3)
    assert cov == 100.0 and f == p


def test_filter_one_pos_not_significant():
    p = Pileup.from_pos_data([[assemble(1, [0, 5, 9], [0, 1, 2], [0, 0, 0])]])
    f, cov = run_filter(p, np.arange(3), 1e-3)
    assert cov == 0 and f.n_chr == 1 and f.n_loci == 0


def test_filter_all_significant():
    pd = one_pos()
    p = Pileup.from_pos_data([[(i + 1, pd[1], pd[2]) for i in range(100)] for _ in range(23)])
    f, cov = run_filter(p, np.arange(100), 1e-3)
    assert cov == 100 and f.n_chr == 23 and f == p


def test_filter_none_significant():
    pd = assemble(1, [0, 5, 9], [1, 3, 5], [0, 0, 0])
    p = Pileup.from_pos_data([[(i + 1, pd[1], pd[2]) for i in range(100)]])
    f, cov = run_filter(p, np.arange(10), 1e-3)
    assert cov == 0 and f.n_chr == 1 and f.n_loci == 0


# ---- golden vectors from the compiled reference ------------------------------------------------------
def test_golden_tuples():
    g = load_golden("filter_tuples")
    for key in g.files:
        if not key.startswith("sig_"):
            continue
        theta = float(key.split("theta")[1].split("_")[0])
        cp = int(key.split("cp")[1])
        got = po.is_significant(g["tuples"], theta, cp)
        assert np.array_equal(got, g[key]), key


@pytest.mark.parametrize("prefix,map_key,cov_key", [("f_", "id_to_pos", "avg_coverage"),
                                                    ("sub_f_", "sub_id_to_pos", "sub_avg_coverage")])
def test_golden_filter_passthrough(prefix, map_key, cov_key):
    g = load_golden("filter_synth")
    f, cov = run_filter(golden_pileup(g), g[map_key], float(g["theta"]))
    assert f == golden_pileup(g, prefix)
    assert cov == float(g[cov_key])


@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_against_compiled_reference():
    rng = np.random.default_rng(0)
    t = rng.integers(0, 60, (5000, 4)).astype(np.uint16)
    t[:, 0] += rng.integers(0, 200, 5000).astype(np.uint16)
    for theta in (0.01, 0.001):
        assert np.array_equal(po.is_significant(t, theta), po.ref_is_significant(t, theta))
