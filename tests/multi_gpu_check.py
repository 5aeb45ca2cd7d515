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
        counts.free()
    # ---- the peer-memory epilogue: no reduction onto one rank; every rank sums the planes of all ranks over its share of
    # the tiles through CUDA IPC mappings, transforms them in the same kernel and writes its share of the matrix into a
    # host matrix shared by the ranks. Without read pairs of order >= 4 it must equal the single-GPU matrix BIT FOR BIT
    # (integer sums, same fp64 expression); with them (fp64 spill planes, added in rank order) to rounding.
    shared = sdist.SharedHostMatrix(ctx, cfg.n_cells)
    cfg_exact = SynthConfig(n_cells=600, coverage=0.3, n_loci=700, n_chr=7, n_clones=3, p_multi=0.02, p_mate=0.05, seed=6)
    for which, pp in (("spill", p), ("exact", make_pileup(cfg_exact)), ("lopsided", p)):
        weights = [int(pp.chr_ptr[c + 1] - pp.chr_ptr[c]) for c in range(pp.n_chr)]
        mine = sdist.partition_chromosomes(weights, world)[rank]
        if which == "lopsided":  # rank 0 holds everything, the others nothing: layouts (planes in use, spill) differ
            mine = list(range(pp.n_chr)) if rank == 0 else []
            which = "spill"
        local_pp = Pileup.concat([pp.loci_range(c, 0, 1 << 40) for c in mine]) if mine else Pileup.empty(0)
        for path in ("gemm", "scatter"):
            f_local, _ = flt.filter_device(local_pp, ident)
            counts = api.Counts(ctx, cfg.n_cells)
            counts.accumulate(f_local, L, ident, eps, h, theta, T, path)
            ep = sdist.SlabEpilogue(counts, device)
            own_sum, peer_sum = ep.checksums()
            assert own_sum == peer_sum, f"checksum of the ranks' planes {own_sum:#x} != checksum of the peer-summed shares {peer_sum:#x}"
            refs = {}
            for norm in ("ADD_MIN", "EXPONENTIATE", "SCALE_MAX_1"):
                shared.array[:] = -7.0
                dist.barrier()
                ep.run(L, eps, h, theta, norm, out_ptr=shared.dev_ptr)
                ctx.synchronize()
                torch.cuda.synchronize()
                dist.barrier()
                if rank == 0:
                    f_all, _ = flt.filter_device(pp, ident)
                    single = api.Counts(ctx, cfg.n_cells)
                    single.accumulate(f_all, L, ident, eps, h, theta, T, path)
                    Ms = single.finalize(L, eps, h, theta, norm)
                    has_spill = single.buffers()[3] > 0
                    assert has_spill == (which == "spill"), "the two pileups must exercise both cases"
                    if has_spill:
                        assert np.abs(shared.array - Ms).max() <= 1e-12 * np.abs(Ms).max(), f"{path}/{norm}: peer-memory epilogue"
                    else:
                        assert np.array_equal(shared.array, Ms, equal_nan=True), f"{path}/{norm}: peer-memory epilogue differs from 1 GPU"
                    refs[norm] = Ms
                    single.free()
                dist.barrier()
            # the share kept on the device: only this rank's tiles (and their mirror images) are written
            dptr = ep.run(L, eps, h, theta, "ADD_MIN")
            ctx.synchronize()
            mine_dev = sdist.tensor_from_ptr(dptr, cfg.n_cells ** 2, torch.float64, device).cpu().numpy().reshape(cfg.n_cells, -1)
            t0_, t1_ = counts.slab_range()
            assert (t0_, t1_) == sdist.slab_tiles(cfg.n_cells, rank, world)
            nb = (cfg.n_cells + 31) // 32
            ref = torch.from_numpy(refs["ADD_MIN"]).to(device) if rank == 0 else torch.empty((cfg.n_cells, cfg.n_cells),
                                                                                           dtype=torch.float64, device=device)
            dist.broadcast(ref, src=0)
            ref = ref.cpu().numpy()
            tol = 1e-12 * np.abs(ref).max() if which == "spill" else 0.0
            for t in range(t0_, t1_, max(1, (t1_ - t0_) // 50)):
                bi, bj = sdist.tri_tile(t, nb)
                blk = (slice(bi * 32, bi * 32 + 32), slice(bj * 32, bj * 32 + 32))
                assert np.abs(mine_dev[blk] - ref[blk]).max() <= tol and np.abs(mine_dev[blk[1], blk[0]] - ref[blk[1], blk[0]]).max() <= tol
            if rank == 0:
                print(f"multi_gpu_check[{path}, peer-memory epilogue, {which}] world={world}: checksums agree, matrix "
                      f"{'within 1e-12 of' if which == 'spill' else 'bit-identical to'} 1 GPU OK", flush=True)
            ep.close()
            counts.free()
            dist.barrier()
    shared.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
