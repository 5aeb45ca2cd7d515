"""Deterministic synthetic pileups (host / numpy) in the shape SURVEY.md §8(d) describes.

Per chromosome, loci sit at increasing positions with a mean spacing of ``spacing`` bp. Each locus
belongs to one class:

* noise         every cell carries the reference base (rejected by the filter unless the pooled
                coverage is extreme);
* germline het  every cell is 50/50 ref/alt (rejected: majority < 1.5 x second, is_significant.cpp:101);
* somatic het   the cells of ONE clone are 50/50 ref/alt, all others carry the reference base.

The number of reads of a cell at a locus is Poisson(coverage); every read base is flipped with
probability ``theta`` to one of the three other bases. Read ids are unique per (cell, fragment)
inside a chromosome. Two knobs exercise the reference's read semantics (similarity_matrix.cpp):
``p_multi`` — probability that a fragment also covers the next locus (if it lies within
``max_fragment_length`` of the fragment start), applied repeatedly; ``p_mate`` — probability that
a (fragment, locus) appears twice (overlapping mates), the second copy disagreeing with
probability ``p_mate_mismatch``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from .pileup import Pileup


@dataclass
class SynthConfig:
    n_cells: int = 100
    coverage: float = 0.1           # reads per cell per locus
    n_loci: int = 2000              # pre-filter loci per chromosome
    n_chr: int = 1
    n_clones: int = 2
    clone_fractions: Optional[tuple] = None   # defaults to equal sizes
    frac_somatic: float = 0.5
    frac_germline: float = 0.1      # remainder is noise
    theta: float = 0.01             # sequencing error used to corrupt bases
    spacing: int = 400
    max_fragment_length: int = 1000
    p_multi: float = 0.0
    p_mate: float = 0.0
    p_mate_mismatch: float = 0.2
    seed: int = 1
    clone_of_cell: np.ndarray = field(default=None, repr=False)


def clone_assignment(cfg: SynthConfig) -> np.ndarray:
    fr = cfg.clone_fractions or tuple([1.0 / cfg.n_clones] * cfg.n_clones)
    bounds = np.floor(np.cumsum(fr) / np.sum(fr) * cfg.n_cells + 1e-9).astype(np.int64)
    clone = np.zeros(cfg.n_cells, np.int64)
    lo = 0
    for k, hi in enumerate(bounds):
        clone[lo:hi] = k
        lo = hi
    return clone


def make_pileup(cfg: SynthConfig) -> Pileup:
    rng = np.random.default_rng(cfg.seed)
    clone = clone_assignment(cfg) if cfg.clone_of_cell is None else np.asarray(cfg.clone_of_cell)
    parts = []
    for c in range(cfg.n_chr):
        parts.append(_make_chromosome(cfg, clone, rng))
    return Pileup.concat(parts)


def _bases_for(rng, locus, cell, ref, alt, cls, som_clone, clone, theta):
    """Draw the observed base of reads (locus[i], cell[i])."""
    n = locus.size
    het = (cls[locus] == 1) | ((cls[locus] == 2) & (clone[cell] == som_clone[locus]))
    take_alt = het & (rng.random(n) < 0.5)
    base = np.where(take_alt, alt[locus], ref[locus])
    err = rng.random(n) < theta
    base = np.where(err, (base + rng.integers(1, 4, n)) & 3, base)
    return base.astype(np.uint16)


def _make_chromosome(cfg: SynthConfig, clone: np.ndarray, rng) -> Pileup:
    P, N = cfg.n_loci, cfg.n_cells
    gaps = rng.integers(1, 2 * cfg.spacing, P)
    position = (1000 + np.cumsum(gaps)).astype(np.int64)
    u = rng.random(P)
    cls = np.where(u < cfg.frac_somatic, 2, np.where(u < cfg.frac_somatic + cfg.frac_germline, 1, 0))
    ref = rng.integers(0, 4, P)
    alt = (ref + rng.integers(1, 4, P)) & 3
    som_clone = rng.integers(0, max(1, int(clone.max()) + 1), P)

    # fragments: one per (locus, read); count per locus ~ Poisson(N * coverage), cells uniform
    n_new = rng.poisson(N * cfg.coverage, P)
    locus = np.repeat(np.arange(P), n_new)
    cell = rng.integers(0, N, locus.size)
    frag = np.arange(locus.size, dtype=np.int64)           # fragment (read) id
    start = position[locus]
    e_locus, e_cell, e_frag = [locus], [cell], [frag]
    # extend fragments over following loci
    cur_l, cur_c, cur_f, cur_s = locus, cell, frag, start
    while cfg.p_multi > 0 and cur_l.size:
        nxt = cur_l + 1
        ok = (nxt < P) & (rng.random(cur_l.size) < cfg.p_multi)
        ok[ok] &= position[nxt[ok]] - cur_s[ok] < cfg.max_fragment_length
        cur_l, cur_c, cur_f, cur_s = nxt[ok], cur_c[ok], cur_f[ok], cur_s[ok]
        e_locus.append(cur_l); e_cell.append(cur_c); e_frag.append(cur_f)
    locus = np.concatenate(e_locus); cell = np.concatenate(e_cell); frag = np.concatenate(e_frag)
    base = _bases_for(rng, locus, cell, ref, alt, cls, som_clone, clone, cfg.theta)
    # overlapping mates: a second entry of the same fragment at the same locus
    if cfg.p_mate > 0:
        dup = rng.random(locus.size) < cfg.p_mate
        dbase = base[dup].copy()
        mism = rng.random(dbase.size) < cfg.p_mate_mismatch
        dbase[mism] = (dbase[mism] + rng.integers(1, 4, int(mism.sum())).astype(np.uint16)) & 3
        locus = np.concatenate([locus, locus[dup]]); cell = np.concatenate([cell, cell[dup]])
        frag = np.concatenate([frag, frag[dup]]); base = np.concatenate([base, dbase])
    # random entry order inside each locus
    order = np.lexsort((rng.random(locus.size), locus))
    locus, cell, frag, base = locus[order], cell[order], frag[order], base[order]
    # read ids: a random injective relabelling, so ids are neither sorted nor dense
    relabel = rng.permutation(int(frag.max()) + 1 if frag.size else 1).astype(np.uint32)
    read_id = relabel[frag] * np.uint32(3) + np.uint32(7)
    counts = np.bincount(locus, minlength=P)
    keep = counts > 0                                       # a pileup never holds empty loci
    row_ptr = np.concatenate([[0], np.cumsum(counts[keep])]).astype(np.uint64)
    gb_t = np.uint32 if cfg.n_cells > 16383 else np.uint16  # beyond the reference's 14-bit group ids: a wide pileup
    gid_base = ((cell.astype(gb_t) << 2) | base.astype(gb_t)).astype(gb_t)
    n_loci = int(keep.sum())
    return Pileup(np.array([0, n_loci], np.uint64), row_ptr, position[keep].astype(np.uint32), read_id, gid_base)
