"""Phases of the cross-rank reduction at bench shape (8000 cells, one bench batch per rank), 2+ ranks:
   python -m torch.distributed.run --nproc-per-node 2 profiles/reduce_probe.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from secedo_b200 import api  # noqa: E402
from secedo_b200 import dist as sdist  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
ctx = api.Context(local)
stream = torch.cuda.Stream(device)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
N = 8000
dev = ctx.synth_pileup(N, 0.5, 2, 32768, n_clones=2, theta=0.001, p_multi=0.005, p_mate=0.01, seed=1000 + rank)
ident = np.arange(N, dtype=np.uint32)
f, _ = api.Filter(0.001, 4, ctx).filter_device(dev, ident)
c = api.Counts(ctx, N)
st = c.accumulate(f, 1000, ident, 0.01, 0.15, 0.001, 8, "gemm")


def timed(fn, n=5):
    best = 1e9
    for _ in range(n):
        dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t) * 1e3)
    return best


res = {}
for mode in ("0", "1"):
    os.environ["SECEDO_B200_SPARSE_REDUCE"] = mode
    res["reduce_counts sparse=" + mode] = timed(lambda: sdist.reduce_counts(c, device, dst=0))
res["sparse_pack"] = timed(lambda: c.sparse_pack(2))
ip, vp, nnz = c.sparse_pack(2)
res["pack_range(0,2)"] = timed(lambda: c.pack_range(0, 2))
res["pack all"] = timed(lambda: c.pack())
ptr, n = c.pack_range(0, 2)
t2 = sdist.tensor_from_ptr(ptr, n, torch.int32, device)
res["nccl reduce 2 planes"] = timed(lambda: dist.reduce(t2, dst=0))
ptr, n = c.pack()
t5 = sdist.tensor_from_ptr(ptr, n, torch.int32, device)
res["nccl reduce all planes"] = timed(lambda: dist.reduce(t5, dst=0))
idx_t, val_t = sdist.tensor_from_ptr(ip, nnz, torch.int32, device), sdist.tensor_from_ptr(vp, nnz, torch.int32, device)
res["exchange_sparse"] = timed(lambda: sdist.exchange_sparse(idx_t, val_t, 0))
if rank == 0:
    print(f"world {world}, planes in use {c.buffers()[1] // (N * N)}, nnz of the sparse planes on rank 0: {nnz} "
          f"({8 * nnz / 1e6:.0f} MB as lists, {4 * 3 * N * (N - 1) // 2 / 1e6:.0f} MB packed dense), pairs {st['n_pairs_multi']}")
    for k, v in res.items():
        print(f"{k:28s} {v:8.3f} ms")
dist.barrier()
dist.destroy_process_group()
