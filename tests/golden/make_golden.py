"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled into oracle/_ref
(oracle/Makefile). Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

Every fixture stores the input pileup (CSR), the parameters and the reference's outputs, so the
tests never need /root/reference at run time. Nothing here is imported by the product.
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pyoracle as po  # noqa: E402
from secedo_b200.pileup import NO_POS, Pileup  # noqa: E402
from secedo_b200.synth import SynthConfig, make_pileup  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SECEDO_REF", "/root/reference")


def dense_cell_map(p, n_groups=10000):
    """group_id_to_pos that maps the cells present in p to 0..n-1 and everything else to NO_POS."""
    cells = np.unique(p.gid_base >> 2)
    m = np.full(n_groups, NO_POS, np.uint32)
    m[cells] = np.arange(cells.size, dtype=np.uint32)
    return m, int(cells.size)


def reference_style_pileup(num_cells=100, num_pos=1200, avg_coverage=0.2, seq_error_rate=0.05, seed=5):
    """The generator of the reference's own end-to-end test (tests/test_spectral_clustering.cpp:
    210-250): consecutive positions, all read ids distinct, two groups of cells that differ at
    the 'significant' positions."""
    rng = np.random.default_rng(seed)
    hi = int(2 * avg_coverage * num_cells)
    chrom, rid = [], 0
    for pos in range(num_pos):
        coverage = rng.integers(1, hi + 1)
        significant = rng.random() < 0.5
        ids, gbs = [], []
        for cell in range(num_cells):
            if rng.integers(1, hi + 1) <= coverage:
                base = (0 if cell < num_cells // 2 else 1) if significant else 2
                if rng.random() < seq_error_rate:
                    base = int(rng.integers(1, hi + 1)) % 4
                gbs.append(cell << 2 | base)
                ids.append(rid)
                rid += 1
        chrom.append((pos, ids, gbs))
    return Pileup.from_pos_data([chrom])


def save_similarity_case(name, p, num_cells, L, gmap, eps, h, theta, threads, norms=("ADD_MIN",)):
    out = dict(chr_ptr=p.chr_ptr, row_ptr=p.row_ptr, position=p.position, read_id=p.read_id, gid_base=p.gid_base,
               num_cells=num_cells, L=L, gmap=gmap, eps=eps, h=h, theta=theta, threads=np.array(threads),
               norms=np.array(norms))
    for t in threads:
        for norm in norms:
            M, _ = po.ref_similarity(p, num_cells, L, gmap, eps, h, theta, t, norm)
            out[f"M_t{t}_{norm}"] = M
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loci", p.n_loci, "entries", p.n_entries, "cells", num_cells)


def span_chain_pileup():
    """A read id chained over >= max_fragment_length: the same id at nine consecutive loci ~300 bp apart (a long-insert
    pair with kept loci in between). The reference retires the read once start + L <= position and starts a new one at
    the next entry with that id (similarity_matrix.cpp:348-372, 379-382); the pileup is dense enough (>= 4 * num_threads
    new reads per locus) for the batch to fire at every locus, so its result does not depend on batch timing. Returns
    (chained pileup, the same pileup with the chain relabelled into the reads the reference makes of it)."""
    cfg = SynthConfig(n_cells=12, coverage=1.0, n_loci=60, n_chr=1, spacing=300, p_multi=0.2, p_mate=0.05, theta=0.05, seed=3)
    p = make_pileup(cfg)
    pos = p.position.astype(np.int64)
    rid, gb = p.read_id.copy(), p.gid_base.copy()
    relabel = rid.copy()
    next_id, seg_start, cur = 4_000_000_100, None, None
    for l in range(10, 19):
        e = int(p.row_ptr[l])
        rid[e] = np.uint32(4_000_000_001)
        gb[e] = (3 << 2) | (gb[e] & 3)
        if seg_start is None or pos[l] - seg_start >= 1000:
            seg_start, cur, next_id = pos[l], np.uint32(next_id), next_id + 1
        relabel[e] = cur
    return (Pileup(p.chr_ptr, p.row_ptr, p.position, rid, gb), Pileup(p.chr_ptr, p.row_ptr, p.position, relabel, gb))


def em_cases():
    """(name, filtered pileup, id_to_pos, theta, initial prob_cluster_b) for the EM refinement"""
    cases = []
    cfg = SynthConfig(n_cells=120, coverage=0.5, n_loci=700, n_chr=3, n_clones=2, p_multi=0.1, p_mate=0.05, theta=0.01, seed=41)
    p = make_pileup(cfg)
    ident = np.arange(cfg.n_cells, dtype=np.uint32)
    rf, _, _ = po.ref_filter(p, ident, 0.01, 4, 1)
    f = Pileup(rf.chr_ptr, rf.row_ptr, rf.position, rf.read_id, rf.gid_base)
    rng = np.random.default_rng(9)
    # a spectral-clustering-like start: hard 0/1 labels along the first half / second half, 15 % of them wrong
    start = (np.arange(cfg.n_cells) >= cfg.n_cells // 2).astype(np.float64)
    flip = rng.random(cfg.n_cells) < 0.15
    start[flip] = 1 - start[flip]
    cases.append(("hard_labels", f, ident, 0.01, start))
    cases.append(("soft_labels", f, ident, 0.001, np.clip(start * 0.6 + 0.2 + rng.normal(0, 0.05, cfg.n_cells), 0.01, 0.99)))
    # the group-id / id_to_pos asymmetry of the reference (expectation_maximization.cpp:24 vs :88-89): cells 0..59
    # only, the likelihoods accumulated at a permuted position
    sub = np.full(cfg.n_cells, NO_POS, np.uint32)
    sub[:60] = rng.permutation(60).astype(np.uint32)
    rf, _, _ = po.ref_filter(p, sub, 0.01, 4, 1)
    fs = Pileup(rf.chr_ptr, rf.row_ptr, rf.position, rf.read_id, rf.gid_base)
    cases.append(("permuted_subcluster", fs, sub, 0.01, rng.uniform(0.3, 0.7, 60)))
    return cases


def make_em():
    out = {"names": []}
    for name, f, m, theta, start in em_cases():
        final, _ = po.ref_expectation_maximization(f, m, theta, start)
        mine, it = po.expectation_maximization(f, m, theta, start)
        assert np.array_equal(final, mine), name  # the restatement follows the reference's summation order exactly
        out["names"].append(name)
        for k, v in dict(chr_ptr=f.chr_ptr, row_ptr=f.row_ptr, position=f.position, read_id=f.read_id, gid_base=f.gid_base,
                         id_to_pos=m, theta=theta, start=start, final=final, iterations=it).items():
            out[f"{name}_{k}"] = v
        print("em", name, "loci", f.n_loci, "entries", f.n_entries, "iterations", it, "in b:", int((final > 0.5).sum()),
              "undecided:", int(((final > 0.05) & (final < 0.95)).sum()))
    out["names"] = np.array(out["names"])
    np.savez_compressed(os.path.join(OUT, "em.npz"), **out)


def main():
    assert po.have_ref(), "build oracle/_ref first: make -C oracle ref"
    if "--only-em" in sys.argv:
        return make_em()
    if "--only-span" in sys.argv:
        chained, split = span_chain_pileup()
        save_similarity_case("sim_span_chain", chained, 12, 1000, np.arange(12, dtype=np.uint32), 0.01, 0.5, 0.05, [1, 2])
        g = dict(np.load(os.path.join(OUT, "sim_span_chain.npz")))
        g["split_read_id"] = split.read_id
        np.savez_compressed(os.path.join(OUT, "sim_span_chain.npz"), **g)
        return
    make_em()
    tmp = tempfile.mkdtemp()

    # ---- the reference's text fixtures, read by the reference's own reader ------------------------
    for fname in ("ten_rows.pileup", "six_cells.pileup", "three_rows.pileup"):
        src = os.path.join(tmp, fname)
        shutil.copy(os.path.join(REF, "tests", "data", fname), src)  # the reader writes <file>.bin next to it
        raw, max_len = po.ref_read_pileup_text(src)
        p = Pileup(raw.chr_ptr, raw.row_ptr, raw.position, raw.read_id, raw.gid_base)
        gmap, n = dense_cell_map(p)
        # span == max_len would retire a read while it still receives bases (outcome depends on batch
        # timing in the reference); the .bin path always uses 1000 (util/pileup_reader.cpp:256)
        L = 1000 if max_len < 1000 else max_len + 1
        save_similarity_case("sim_" + fname.split(".")[0], p, n, L, gmap, 0.01, 0.5, 0.01, [1, 2, 8],
                             ("ADD_MIN", "EXPONENTIATE", "SCALE_MAX_1"))

    # ---- the reference's end-to-end test style (distinct read ids, consecutive positions) ----------
    p = reference_style_pileup()
    ident = np.arange(100, dtype=np.uint32)
    rf, cov, _ = po.ref_filter(p, ident, 0.05, 4, 4)
    f = Pileup(rf.chr_ptr, rf.row_ptr, rf.position, rf.read_id, rf.gid_base)
    save_similarity_case("sim_reference_test_style", f, 100, 500, ident, 0.01, 0.5, 0.05, [4])

    # ---- synthetic, multi-locus reads + overlapping mates, 3 chromosomes -----------------------------
    cfg = SynthConfig(n_cells=40, coverage=0.3, n_loci=500, n_chr=3, p_multi=0.4, p_mate=0.15, theta=0.02, seed=11)
    p = make_pileup(cfg)
    ident = np.arange(cfg.n_cells, dtype=np.uint32)
    rf, cov, _ = po.ref_filter(p, ident, 0.01, 4, 1)
    f = Pileup(rf.chr_ptr, rf.row_ptr, rf.position, rf.read_id, rf.gid_base)
    save_similarity_case("sim_synth_multi", f, cfg.n_cells, 1000, ident, 0.01, 0.5, 0.01, [1, 2, 3, 8],
                         ("ADD_MIN", "EXPONENTIATE", "SCALE_MAX_1"))
    # ---- a read id chained over >= L bp (a drop-in run must complete such pileups like the reference) ----
    chained, split = span_chain_pileup()
    save_similarity_case("sim_span_chain", chained, 12, 1000, np.arange(12, dtype=np.uint32), 0.01, 0.5, 0.05, [1, 2])
    g = dict(np.load(os.path.join(OUT, "sim_span_chain.npz")))
    g["split_read_id"] = split.read_id
    np.savez_compressed(os.path.join(OUT, "sim_span_chain.npz"), **g)
    # sub-cluster: only the first clone's cells take part (id_to_pos with NO_POS)
    sub = np.full(cfg.n_cells, NO_POS, np.uint32)
    sub[:20] = np.arange(20)
    rf, cov_sub, _ = po.ref_filter(p, sub, 0.01, 4, 1)
    fs = Pileup(rf.chr_ptr, rf.row_ptr, rf.position, rf.read_id, rf.gid_base)
    save_similarity_case("sim_synth_subcluster", fs, 20, 1000, sub, 0.01, 0.15, 0.001, [2])

    # ---- filter: pass-through of a whole pileup + decisions on count tuples ---------------------------
    rf, cov, _ = po.ref_filter(p, ident, 0.01, 4, 1)
    np.savez_compressed(os.path.join(OUT, "filter_synth.npz"), chr_ptr=p.chr_ptr, row_ptr=p.row_ptr,
                        position=p.position, read_id=p.read_id, gid_base=p.gid_base, id_to_pos=ident, theta=0.01,
                        f_chr_ptr=rf.chr_ptr, f_row_ptr=rf.row_ptr, f_position=rf.position, f_read_id=rf.read_id,
                        f_gid_base=rf.gid_base, avg_coverage=cov,
                        sub_id_to_pos=sub, sub_f_chr_ptr=fs.chr_ptr, sub_f_row_ptr=fs.row_ptr,
                        sub_f_position=fs.position, sub_f_read_id=fs.read_id, sub_f_gid_base=fs.gid_base,
                        sub_avg_coverage=cov_sub)
    rng = np.random.default_rng(3)
    tuples = []
    for cov_hi in (12, 40, 120, 260, 1000, 5000, 30000):
        n = 3000
        total = rng.integers(2, cov_hi, n)
        frac = rng.random((n, 3)) * np.array([0.45, 0.08, 0.03])
        minor = np.floor(frac * total[:, None]).astype(np.int64)
        major = total - minor.sum(1)
        t = np.concatenate([major[:, None], minor], 1)
        for row in t:
            rng.shuffle(row)
        tuples.append(t)
    tuples = np.clip(np.concatenate(tuples), 0, 65535).astype(np.uint16)
    dec = {}
    for theta in (0.01, 0.001, 0.05):
        for cp in (0, 2, 4):
            dec[f"sig_theta{theta}_cp{cp}"] = po.ref_is_significant(tuples, theta, cp)
    np.savez_compressed(os.path.join(OUT, "filter_tuples.npz"), tuples=tuples, **dec)
    print("filter tuples", tuples.shape, {k: int(v.sum()) for k, v in dec.items()})

    # ---- LS / LD tables ------------------------------------------------------------------------------
    tabs = {}
    for i, (e, h, t, L) in enumerate([(0.01, 0.5, 0.01, 1000), (0.01, 0.5, 0.001, 1000), (0.01, 0.15, 0.001, 1000),
                                      (0.01, 0.5, 0.05, 500), (0.02, 0.3, 0.01, 12)]):
        n = min(16, L)
        ls, ld = po.ref_log_probs(e, h, t, L, n if L > 20 else 6)
        tabs[f"params{i}"] = np.array([e, h, t, L])
        tabs[f"ls{i}"] = ls
        tabs[f"ld{i}"] = ld
    np.savez_compressed(os.path.join(OUT, "log_probs.npz"), **tabs)

    # ---- binary pileup ingestion: read_pileup_bin of the reference on files in its own format -----------
    bins = {}
    # (a) the reference's own text fixture, converted with its grouping file (tests/data/six_cells.pileup.group)
    cfg = SynthConfig(n_cells=300, coverage=0.15, n_loci=400, n_chr=2, p_multi=0.1, p_mate=0.05, seed=81)
    p = make_pileup(cfg)
    rng = np.random.default_rng(3)
    grouping = (np.arange(10000) // 3).astype(np.uint16)           # --merge_count 3
    scrambled = rng.permutation(10000).astype(np.uint16) % 117      # a --merge_file style mapping
    for c in range(p.n_chr):
        path = os.path.join(tmp, f"chr{c}.bin")
        raw = p.to_bin(c)
        open(path, "wb").write(raw)
        bins[f"file{c}"] = np.frombuffer(raw, np.uint8)
        chrom = p.loci_range(c, 0, 1 << 40)
        some = np.sort(rng.choice(chrom.position, chrom.position.size // 3, replace=False)).astype(np.uint32)
        some = np.unique(np.concatenate([some, some[:5] + 1]))     # a few positions that are not in the file
        cases = {"ident": (np.arange(10000, dtype=np.uint16), 100, np.zeros(0, np.uint32)),
                 "merge3_cov40": (grouping, 40, np.zeros(0, np.uint32)),
                 "mapfile_pos": (scrambled, 100, some),
                 "pos_early_stop": (np.arange(10000, dtype=np.uint16), 100, some[: some.size // 4])}
        for name, (g, maxcov, pos) in cases.items():
            r, n_cells, max_len = po.ref_read_pileup(path, g, maxcov, pos)
            k = f"c{c}_{name}_"
            bins[k + "grouping"], bins[k + "max_coverage"], bins[k + "positions"] = g, np.int64(maxcov), pos
            bins[k + "n_cells"], bins[k + "max_len"] = np.int64(n_cells), np.int64(max_len)
            for f in ("row_ptr", "position", "read_id", "gid_base"):
                bins[k + f] = getattr(r, f)
            print("bin", c, name, r.n_loci, r.n_entries, n_cells, max_len)
    np.savez_compressed(os.path.join(OUT, "pileup_bin.npz"), **bins)
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
