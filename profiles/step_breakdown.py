"""Wall-clock breakdown of one bench step per rank (debug aid): python [-m torch.distributed.run ...] profiles/step_breakdown.py"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from secedo_b200 import api, dist as sdist
w = bench.WORKLOAD
world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); device = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=device)
ctx = api.Context(lr); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
N = w["n_cells"]; ident = np.arange(N, dtype=np.uint32)
raw = ctx.synth_pileup(N, w["coverage"], w["n_chr"], w["loci_per_chr"], theta=w["theta"], p_multi=w["p_multi"], p_mate=w["p_mate"], seed=1000 + rank)
flt = api.Filter(w["theta"], 4, ctx); counts = api.Counts(ctx, N)
def T(name, fn, acc):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); acc[name] = acc.get(name, 0) + (time.perf_counter() - t0) * 1e3; return r
for it in range(4):
    acc = {}
    f, _ = T("filter", lambda: flt.filter_device(raw, ident), acc)
    T("zero", counts.zero, acc)
    T("accumulate", lambda: counts.accumulate(f, w["L"], ident, w["eps"], w["h"], w["theta"], 8, "auto"), acc)
    T("free", f.free, acc)
    T("reduce", lambda: sdist.reduce_counts(counts, device, 0), acc)
    if rank == 0:
        T("finalize", lambda: counts.finalize(w["L"], w["eps"], w["h"], w["theta"], "ADD_MIN", to_host=False), acc)
    print(rank, it, {k: round(v, 2) for k, v in acc.items()}, flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
