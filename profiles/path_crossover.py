"""First-order read-pair counts by both paths (per-locus pair scatter with int32 atomics vs. int8 tcgen05 GEMM) on the
shapes BASELINE.json names, plus ultra-sparse ones, to justify choose_path (csrc/abi.cu) with measured numbers.
Device-generated pileups of 16 384 pre-filter loci per case; ms are sgpu_stats.ms_first_order (CUDA events), best of 3.
Usage: python profiles/path_crossover.py > gpurun_out/path_crossover.txt"""
import sys

import numpy as np

sys.path.insert(0, ".")
from secedo_b200 import api  # noqa: E402

CASES = [  # name, cells, coverage, theta
    ("cfg1 500 cells 0.05x", 500, 0.05, 0.01),
    ("cfg2 2000 cells 0.1x", 2000, 0.1, 0.01),
    ("cfg4 10000 cells 0.05x (breast)", 10000, 0.05, 0.001),
    ("cfg3 8000 cells 0.5x", 8000, 0.5, 0.001),
    ("8000 cells 0.02x", 8000, 0.02, 0.001),
    ("8000 cells 0.005x", 8000, 0.005, 0.001),
    ("8000 cells 0.002x", 8000, 0.002, 0.001),
    ("16000 cells 0.01x", 16000, 0.01, 0.001),
    ("cfg5 20000 cells 1x (wide ids, 2048 loci)", 20000, 1.0, 0.001),
]
ctx = api.Context(0)
print(f"{'case':34s} {'sig loci':>8s} {'entries':>10s} {'reads/locus':>11s} {'pairs':>13s} {'scatter ms':>10s} {'gemm ms':>8s} "
      f"{'Gpairs/s':>8s} {'auto':>8s}")
for name, n, cov, theta in CASES:
    dev = ctx.synth_pileup(n, cov, 1, 2048 if n >= 20000 else 16384, n_clones=4, theta=theta, frac_somatic=0.5, frac_germline=0.0, p_multi=0.005,
                           p_mate=0.01, seed=7)
    ident = np.arange(n, dtype=np.uint32)
    # keep every locus (the filter would drop most ultra-sparse ones): density is what is being measured
    f = dev
    c = api.Counts(ctx, n)
    res = {}
    for path in ("scatter", "gemm", "auto"):
        best, st = 1e9, None
        if path == "scatter" and n >= 20000:  # 4e11 pairs: minutes; the GEMM path is the only sensible one there
            res[path] = (float("nan"), {"n_pairs_first": 0})
            continue
        for _ in range(3):
            c.zero()
            st = c.accumulate(f, 1000, ident, 0.01, 0.15, theta, 8, path)
            best = min(best, st["ms_first_order"])
        res[path] = (best, st)
    pairs = res["scatter"][1]["n_pairs_first"]
    L, E = f.n_loci, f.n_entries
    mark = "" if res["auto"][1]["path_used"] == ("scatter" if res["scatter"][0] < res["gemm"][0] else "gemm") else "  <-- auto picked the slower path"
    print(f"{name:34s} {L:8d} {E:10d} {E / max(L, 1):11.1f} {pairs:13d} {res['scatter'][0]:10.3f} {res['gemm'][0]:8.3f} "
          f"{pairs / res['scatter'][0] / 1e6:8.2f} {res['auto'][1]['path_used']:>8s}{mark}", flush=True)
    c.free()
    dev.free()
