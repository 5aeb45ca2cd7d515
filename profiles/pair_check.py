"""Quick check of the CTA-pair GEMM against the scatter path (debug aid; run under `timeout`)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from secedo_b200 import api
n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 700
ctx = api.Context(0)
dev = ctx.synth_pileup(n_cells, 0.4, 2, 300, n_clones=3, theta=0.001, p_multi=0.05, p_mate=0.03, seed=3)
ident = np.arange(n_cells, dtype=np.uint32)
fdev, _ = api.Filter(0.001, 4, ctx).filter_device(dev, ident)
res = {}
for path in ("scatter", "gemm"):
    c = api.Counts(ctx, n_cells)
    c.accumulate(fdev, 1000, ident, 0.01, 0.5, 0.001, 8, path)
    res[path] = c.download()
    c.free()
ok = all(np.array_equal(a, b) for a, b in zip(res["gemm"], res["scatter"]))
print("pairs =", os.environ.get("SECEDO_B200_GEMM_PAIRS"), "n_cells", n_cells, "loci", fdev.n_loci, "MATCH" if ok else "MISMATCH",
      int(res["gemm"][0].sum()), int(res["scatter"][0].sum()))
