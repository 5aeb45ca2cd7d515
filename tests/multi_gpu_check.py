"""Multi-GPU parity check, launched by tests/test_gpu_multi.py (or by hand) as
    python -m torch.distributed.run --nproc-per-node N tests/multi_gpu_check.py
Every rank filters + accumulates ITS chromosomes (secedo_b200.dist.partition_chromosomes), the count
planes are summed with one NCCL reduce, rank 0 finalizes. Rank 0 then recomputes everything on its own
GPU alone and with the oracle: integer counts must be bit-identical, the matrix must be identical to
the single-GPU one and within 1e-6 * max|M| of the oracle's."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import pyoracle as po  # noqa: E402
from secedo_b200 import api  # noqa: E402
from secedo_b200 import dist as sdist  # noqa: E402
from secedo_b200.pileup import Pileup  # noqa: E402
from secedo_b200.synth import SynthConfig, make_pileup  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    ctx = api.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    cfg = SynthConfig(n_cells=600, coverage=0.3, n_loci=700, n_chr=7, n_clones=3, p_multi=0.1, p_mate=0.05, seed=5)
    p = make_pileup(cfg)                                   # identical on every rank
    ident = np.arange(cfg.n_cells, dtype=np.uint32)
    L, eps, h, theta, T = 1000, 0.01, 0.5, 0.01, 8
    weights = [int(p.chr_ptr[c + 1] - p.chr_ptr[c]) for c in range(p.n_chr)]
    mine = sdist.partition_chromosomes(weights, world)[rank]
    local_p = Pileup.concat([p.loci_range(c, 0, 1 << 40) for c in mine]) if mine else Pileup.empty(0)
    flt = api.Filter(theta, 4, ctx)
    for path, sparse in (("gemm", "1"), ("gemm", "0"), ("scatter", None)):
        # second-order planes as lists of non-zeros, as dense planes, and whichever the size estimate picks
        if sparse is None:
            os.environ.pop("SECEDO_B200_SPARSE_REDUCE", None)
        else:
            os.environ["SECEDO_B200_SPARSE_REDUCE"] = sparse
        f_local, _ = flt.filter_device(local_p, ident)
        counts = api.Counts(ctx, cfg.n_cells)
        # rank 1 accumulates without multi-locus reads knowledge of the others: layouts get reconciled
        counts.accumulate(f_local, L, ident, eps, h, theta, T, path)
        t0 = time.perf_counter()
        sdist.reduce_counts(counts, device, dst=0)
        torch.cuda.synchronize()
        t_red = time.perf_counter() - t0
        if rank == 0:
            S1, D1, H, hist = counts.download()
            M = counts.finalize(L, eps, h, theta, "ADD_MIN")
            f_all, _ = flt.filter_device(p, ident)
            single = api.Counts(ctx, cfg.n_cells)
            single.accumulate(f_all, L, ident, eps, h, theta, T, path)
            s = single.download()
            Ms = single.finalize(L, eps, h, theta, "ADD_MIN")
            for a, b, name in zip((S1, D1, H, hist), s, ("S1", "D1", "H", "hist")):
                assert np.array_equal(a, b), f"{path}: {name} differs between {world} GPUs and 1 GPU"
            assert np.array_equal(M, Ms), f"{path}: matrix differs between {world} GPUs and 1 GPU"
            o = po.similarity(f_all.download(), cfg.n_cells, L, ident, eps, h, theta, T, "ADD_MIN")
            assert np.array_equal(S1, o.S1) and np.array_equal(D1, o.D1) and np.array_equal(H, o.H)
            assert np.abs(M - o.M).max() <= 1e-6 * np.abs(o.M).max()
            print(f"multi_gpu_check[{path}, sparse={sparse}] world={world}: counts bit-identical to 1 GPU and to the oracle, "
                  f"reduce {t_red * 1e3:.2f} ms OK", flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
