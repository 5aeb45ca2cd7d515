"""The oracle's restatement of expectation_maximization() (oracle/secedo_oracle.c) against the reference's own
known-answer tests (tests/test_expectation_maximization.cpp, replayed), the golden vectors generated from the compiled
reference (tests/golden/em.npz) and, where oracle/_ref is present, the compiled reference itself (bit for bit: the
restatement keeps the reference's summation order)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import pyoracle as po
from secedo_b200.pileup import Pileup

THETA = 1e-3


def repeated_locus(gid_base, n_loci=3):
    """the PosData `one_pos` of the reference's tests, n_loci times in one chromosome"""
    gb = np.array(gid_base, np.uint16)
    k = gb.size
    return Pileup(np.array([0, n_loci], np.uint64), (np.arange(n_loci + 1) * k).astype(np.uint64),
                  np.full(n_loci, 1234, np.uint32), np.tile(np.arange(1000, 1000 + k, dtype=np.uint32), n_loci),
                  np.tile(gb, n_loci))


# (cells << 2 | base at the locus, id_to_pos, start, check) - tests/test_expectation_maximization.cpp:15-96
REFERENCE_TESTS = {
    "OneCell": ([], [], [1.0], lambda p: p[0] == 1.0),
    "TwoCellsSame": ([0 << 2 | 1, 1 << 2 | 1], [0, 1], [0.3, 0.4], lambda p: abs(p[1] - p[0]) < 1e-3),
    "TwoCellsDifferent": ([0 << 2 | 1, 1 << 2 | 2], [0, 1], [0.01, 0.02], lambda p: abs(abs(p[0] - p[1]) - 1.0) < 1e-3),
    "FourCellsTwoGroups22": ([0 << 2 | 1, 1 << 2 | 1, 2 << 2 | 2, 3 << 2 | 2], [0, 1, 2, 3], [0.9, 0.02, 0.03, 0.9],
                             lambda p: np.abs(p - [0, 0, 1, 1]).max() < 1e-3),
    "FourCellsTwoGroups31": ([0 << 2 | 2, 1 << 2 | 1, 2 << 2 | 2, 3 << 2 | 2], [0, 1, 2, 3], [0.9, 0.9, 0.03, 0.1],
                             lambda p: np.abs(p - [0, 1, 0, 0]).max() < 1e-3),
}


def reference_test_input(name):
    gb, m, start, check = REFERENCE_TESTS[name]
    p = repeated_locus(gb) if gb else Pileup(np.array([0, 0], np.uint64), np.array([0], np.uint64), np.zeros(0, np.uint32),
                                             np.zeros(0, np.uint32), np.zeros(0, np.uint16))
    return p, np.array(m, np.uint32), np.array(start, np.float64), check


@pytest.mark.parametrize("name", sorted(REFERENCE_TESTS))
def test_oracle_reference_known_answers(name):
    p, m, start, check = reference_test_input(name)
    prob, it = po.expectation_maximization(p, m, THETA, start)
    assert check(prob), prob
    if po.have_ref():
        ref, _ = po.ref_expectation_maximization(p, m, THETA, start)
        assert np.array_equal(ref, prob)


def golden_em_cases():
    g = load_golden("em")
    for name in g["names"]:
        name = str(name)
        p = Pileup(*(g[f"{name}_{k}"] for k in ("chr_ptr", "row_ptr", "position", "read_id", "gid_base")))
        yield name, p, g[f"{name}_id_to_pos"], float(g[f"{name}_theta"]), g[f"{name}_start"], g[f"{name}_final"], int(g[f"{name}_iterations"])


def test_oracle_matches_golden_vectors():
    n = 0
    for name, p, m, theta, start, final, iterations in golden_em_cases():
        prob, it = po.expectation_maximization(p, m, theta, start)
        assert np.array_equal(prob, final) and it == iterations, name
        n += 1
    assert n == 3


def test_oracle_rejects_out_of_range_groups():
    p = repeated_locus([5 << 2 | 1, 1 << 2 | 1])
    with pytest.raises(ValueError):
        po.expectation_maximization(p, np.arange(6, dtype=np.uint32), THETA, [0.5, 0.5])  # prob_cluster[5] out of range
