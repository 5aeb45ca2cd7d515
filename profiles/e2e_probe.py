"""Where the end-to-end step spends its time: one chromosome of the bench workload (8000 cells, 0.5x, 32768 loci) from
pinned host memory. Wall clock around synchronised calls, best of 3. Usage: python profiles/e2e_probe.py"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from secedo_b200 import api  # noqa: E402
from secedo_b200.pileup import Pileup  # noqa: E402

N = 8000
ctx = api.Context(0)
dev = ctx.synth_pileup(N, 0.5, 1, 32768, n_clones=2, theta=0.001, p_multi=0.005, p_mate=0.01, seed=1000)
host = dev.download()
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
keep = [pin(getattr(host, k)) for k in ("row_ptr", "position", "read_id", "gid_base")]
hp = Pileup(host.chr_ptr, *(t.numpy() for t in keep))
ident = np.arange(N, dtype=np.uint32)
flt = api.Filter(0.001, 4, ctx)
out = torch.empty((N, N), dtype=torch.float64).pin_memory().numpy()
nbytes = sum(t.numpy().nbytes for t in keep)


def best(fn, n=3):
    ts = []
    for _ in range(n):
        ctx.synchronize()
        t = time.perf_counter()
        r = fn()
        ctx.synchronize()
        ts.append((time.perf_counter() - t) * 1e3)
        if hasattr(r, "free"):
            r.free()
    return min(ts)


def up_full():
    p = ctx.upload_async(hp)
    p.dims()
    ctx._lib.sgpu_pileup_download(ctx._h, p._h, None, None, None, None, None)  # waits for the copies
    return p


def up_lazy():
    p = ctx.upload_lazy_async(hp)
    ctx._lib.sgpu_pileup_download(ctx._h, p._h, None, None, None, None, None)
    return p


print(f"entries {host.n_entries}, bytes {nbytes / 1e6:.0f} MB (read ids {keep[2].numpy().nbytes / 1e6:.0f} MB)")
t = best(up_full)
print(f"upload all arrays (DMA):            {t:7.2f} ms  {nbytes / t / 1e6:6.1f} GB/s")
t = best(up_lazy)
print(f"upload without read ids (DMA):      {t:7.2f} ms  {(nbytes - keep[2].numpy().nbytes) / t / 1e6:6.1f} GB/s")
full, lazy = ctx.upload(hp), ctx.upload_lazy_async(hp)
t1 = best(lambda: flt.filter_device(full, ident)[0])
f, _ = flt.filter_device(full, ident)
t2 = best(lambda: flt.filter_device(lazy, ident)[0])
print(f"filter, resident pileup:            {t1:7.2f} ms")
print(f"filter, read ids pulled (zero copy):{t2:7.2f} ms  -> pull of {4 * f.n_entries / 1e6:.0f} MB at {4 * f.n_entries / (t2 - t1) / 1e6:6.1f} GB/s")
c = api.Counts(ctx, N)


def acc():
    c.zero()
    c.accumulate(f, 1000, ident, 0.01, 0.15, 0.001, 8, "auto")


t = best(acc)
print(f"accumulate:                         {t:7.2f} ms")
t = best(lambda: c.finalize(1000, 0.01, 0.15, 0.001, "ADD_MIN", out=out))
print(f"finalize + D2H of the matrix:       {t:7.2f} ms  {out.nbytes / t / 1e6:6.1f} GB/s")
t = best(lambda: c.finalize(1000, 0.01, 0.15, 0.001, "ADD_MIN", to_host=False))
print(f"finalize, matrix stays in HBM:      {t:7.2f} ms")
