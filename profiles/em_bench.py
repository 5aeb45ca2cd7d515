"""EM refinement on a device-resident filtered pileup of the bench shape (8000 cells, 0.5x, 2 x 32768 loci):
python profiles/em_bench.py"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from secedo_b200 import api  # noqa: E402

N = 8000
ctx = api.Context(0)
dev = ctx.synth_pileup(N, 0.5, 2, 32768, n_clones=2, theta=0.001, p_multi=0.005, p_mate=0.01, seed=1000)
ident = np.arange(N, dtype=np.uint32)
f, _ = api.Filter(0.001, 4, ctx).filter_device(dev, ident)
start = np.random.default_rng(1).uniform(0.3, 0.7, N)
api.expectation_maximization(f, ident, 1, 0.001, start, ctx=ctx, max_iterations=1)
for it in (1, 4, 8):
    _, st = api.expectation_maximization(f, ident, 1, 0.001, start, ctx=ctx, max_iterations=it, return_stats=True)
    print(json.dumps({"entries": f.n_entries, "loci": f.n_loci, **st, "ms_per_iteration": st["ms"] / st["iterations"],
                      "entries_per_s": f.n_entries * st["iterations"] / (st["ms"] * 1e-3)}))
