"""Warm timing of the Laplacian eigen-solver on the device-resident similarity matrix (cfg3 shape: 8000 cells),
next to LAPACK dsyevd (numpy.linalg.eigh, what arma::eig_sym runs in the reference) on the host cores at a size
that finishes in seconds. Usage: python profiles/spectral_bench.py [n_cells] [k] [host_n]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from secedo_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 7
host_n = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
ctx = api.Context(0)
dev = ctx.synth_pileup(n, 0.5, 2, 4096, n_clones=4, theta=0.001, p_multi=0.005, p_mate=0.01, seed=5)
ident = np.arange(n, dtype=np.uint32)
fdev, _ = api.Filter(0.001, 4, ctx).filter_device(dev, ident)
c = api.Counts(ctx, n)
c.accumulate(fdev, 1000, ident, 0.01, 0.15, 0.001, 8, "gemm")
runs = []
for i in range(5):
    t = time.perf_counter()
    ev, vec, st, _ = c.finalize_spectral(1000, 0.01, 0.15, 0.001, "ADD_MIN", k=k, tol=1e-10)
    st["wall_ms"] = (time.perf_counter() - t) * 1e3
    runs.append(st)
    print(json.dumps(st), flush=True)
m = c.finalize(1000, 0.01, 0.15, 0.001, "ADD_MIN")
t = time.perf_counter()
ev2, vec2 = api.spectral_embedding(m, k, 1e-10, ctx=ctx)
host_path_ms = (time.perf_counter() - t) * 1e3
best = min(runs[1:], key=lambda r: r["wall_ms"])
bytes_per_product = 8.0 * n * n
out = {"n_cells": n, "k": k, "eigenvalues": ev.tolist(), "best_warm": best, "from_host_matrix_wall_ms": host_path_ms,
       "block_product_avg_ms": best["ms_matvec"] / best["matvec_launches"],
       "block_product_GBps": bytes_per_product * best["matvec_launches"] / (best["ms_matvec"] * 1e-3) / 1e9}
if host_n:
    sub = np.ascontiguousarray(m[:host_n, :host_n])
    d = sub.sum(1)
    s = 1 / np.sqrt(d)
    lap = np.eye(host_n) - sub * s[:, None] * s[None, :]
    t = time.perf_counter()
    w = np.linalg.eigh(lap)[0]
    dt = time.perf_counter() - t
    out["host_dsyevd"] = {"n": host_n, "seconds": dt, "extrapolated_seconds_at_n_cells": dt * (n / host_n) ** 3}
print(json.dumps(out))
