"""Host-side logic of the multi-GPU path on CPU: chromosome partitioning, layout agreement and the
reduction of count planes, with torch.distributed (gloo, world_size 2). The per-rank counts come from
the oracle on each rank's chromosomes; their reduced sum must equal the oracle on the whole pileup —
the property that makes N-GPU results bit-identical to 1-GPU results."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from secedo_b200.dist import agree_layout, exchange_sparse, partition_chromosomes, reduce_buffers, sparse_pays


def test_partition_is_balanced_and_complete():
    w = [249, 243, 198, 191, 181, 171, 159, 146, 141, 135, 135, 133, 115, 107, 102, 90, 81, 78, 59, 63, 48, 51, 155]
    for world in (1, 2, 4, 8):
        parts = partition_chromosomes(w, world)
        assert sorted(sum(parts, [])) == list(range(len(w)))
        loads = [sum(w[i] for i in p) for p in parts]
        assert max(loads) <= 1.25 * sum(w) / world + 1


def test_sparse_route_decision():
    n, tri = 8000, 8000 * 7999 // 2
    assert sparse_pays([3_000_000] * 8, 0, 3, n)            # 7 x 24 MB of lists against 384 MB of packed planes
    assert not sparse_pays([tri] * 2, 0, 3, n)              # dense planes: lists would be larger
    assert not sparse_pays([0, 0], 0, 0, n)                 # no sparse planes at all
    assert sparse_pays([10 ** 9, 5], 0, 3, n)               # only the OTHER ranks' lists travel


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    from secedo_b200.pileup import Pileup
    from secedo_b200.synth import SynthConfig, make_pileup
    cfg = SynthConfig(n_cells=40, coverage=0.3, n_loci=300, n_chr=5, p_multi=0.4, p_mate=0.1, seed=13)
    p = make_pileup(cfg)
    ident = np.arange(cfg.n_cells, dtype=np.uint32)
    kl, ke, _, _ = po.filter_flags(p, ident, 0.01)
    f = p.select(kl, ke)
    weights = [int(f.chr_ptr[c + 1] - f.chr_ptr[c]) for c in range(f.n_chr)]
    mine = partition_chromosomes(weights, world)[rank]
    local = Pileup.concat([f.loci_range(c, 0, 1 << 40) for c in mine]) if mine else Pileup.empty(0)
    r = po.similarity(local, cfg.n_cells, 1000, ident, 0.01, 0.5, 0.01, 2)
    # rank 1 pretends it saw no multi-locus pairs: layouts differ and must be reconciled
    planes = 9 if rank == 0 else 2
    planes, spill = agree_layout(planes, False)
    assert planes == 9 and spill is False
    bufs = [torch.from_numpy(np.stack([r.S1, r.D1, r.H[0], r.H[1], r.H[2]]).astype(np.int32)),
            torch.from_numpy(r.class_hist.astype(np.int64))]
    reduce_buffers(bufs, dst=0)
    if rank == 0:
        whole = po.similarity(f, cfg.n_cells, 1000, ident, 0.01, 0.5, 0.01, 2)
        ok = (np.array_equal(bufs[0][0].numpy(), whole.S1) and np.array_equal(bufs[0][1].numpy(), whole.D1)
              and np.array_equal(bufs[0][2:].numpy(), whole.H)
              and np.array_equal(bufs[1].numpy().astype(np.uint64), whole.class_hist))
    # the sparse route for the second-order planes: every rank's non-zeros (upper triangle) as (index, value) lists,
    # gathered on rank 0 and added there - what sgpu_counts_sparse_pack / sgpu_counts_sparse_add do on the device
    n = cfg.n_cells
    up = np.triu(np.ones((n, n), bool), 1)
    H = r.H.astype(np.int32) * up
    flat = H.reshape(-1)
    nz = np.flatnonzero(flat)
    idx = torch.from_numpy(nz.astype(np.uint32).view(np.int32).copy())
    val = torch.from_numpy(flat[nz].copy())
    lists = exchange_sparse(idx, val, dst=0)
    if rank == 0:
        acc = flat.copy()
        for li, lv in lists:
            acc[li.numpy().view(np.uint32)] += lv.numpy()
        ok = ok and len(lists) == world - 1 and np.array_equal(acc.reshape(3, n, n), whole.H * up)
        q.put(bool(ok))
    else:
        assert lists == []
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduce_equals_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
