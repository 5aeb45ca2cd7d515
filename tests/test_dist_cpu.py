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
from secedo_b200.dist import (NO_TAIL, agree_layout, exchange_sparse, partition_chromosomes, plan_pieces, reduce_buffers,
                              slab_tiles, sparse_pays, tri_tile, tri_tile_count)


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


def test_pieces_tile_the_chromosomes_with_halos():
    """plan_pieces: the owned ranges tile every chromosome, weights are balanced, and every piece holds all loci less
    than L bp before its first / after its last owned locus (what sgpu_counts_accumulate_range requires)"""
    rng = np.random.default_rng(1)
    L = 1000
    pos = [np.cumsum(rng.integers(1, 800, n)) for n in (1000, 50, 3000, 1, 0, 700)]
    wts = [rng.integers(1, 50, len(p)).astype(float) ** 2 for p in pos]
    for weights in (None, wts):
        for k in (1, 2, 3, 8, 17):
            pieces = plan_pieces(pos, k, L, weights)
            assert len(pieces) == k
            for c, p in enumerate(pos):
                rs = sorted((d["own_lo"], d["own_hi"], d["own_pos_begin"], d["own_pos_end"]) for pc in pieces for d in pc if d["chrom"] == c)
                if len(p) == 0:
                    assert not rs
                    continue
                assert rs[0][0] == 0 and rs[-1][1] == len(p) and rs[0][2] == 0 and rs[-1][3] == NO_TAIL
                assert all(a[1] == b[0] and a[3] == b[2] for a, b in zip(rs, rs[1:]))
                for lo, hi, pb, pe in rs:  # ownership by position selects exactly the planned loci
                    sel = np.flatnonzero((p >= pb) & (p < pe))
                    assert sel[0] == lo and sel[-1] == hi - 1
            for pc in pieces:
                assert len({d["chrom"] for d in pc}) == len(pc), "at most one range per chromosome and piece"
                for d in pc:
                    p = pos[d["chrom"]]
                    need = np.flatnonzero((p > p[d["own_lo"]] - L) & (p < p[d["own_hi"] - 1] + L))
                    assert d["lo"] <= need[0] and d["hi"] >= need[-1] + 1
            ws = wts if weights is not None else [np.ones(len(p)) for p in pos]
            load = [sum(ws[d["chrom"]][d["own_lo"]:d["own_hi"]].sum() for d in pc) for pc in pieces]
            biggest = max(w.max() for w in ws if w.size)
            assert max(load) <= sum(load) / k + biggest + 1e-9


def test_slab_shares_partition_the_upper_triangle():
    """the shares of the peer-memory epilogue: contiguous, disjoint, complete, equal to within one tile; the tile
    numbering is row-major over bj >= bi (what sgpu_slab_raw decodes on the device)"""
    for n in (1, 31, 32, 33, 600, 8000, 20000):
        nb = (n + 31) // 32
        total = tri_tile_count(n)
        assert total == nb * (nb + 1) // 2
        for world in (1, 2, 3, 8):
            shares = [slab_tiles(n, s, world) for s in range(world)]
            assert shares[0][0] == 0 and shares[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(shares, shares[1:]))
            sizes = [b - a for a, b in shares]
            assert max(sizes) - min(sizes) <= 1
        k = 0
        step = max(1, total // 500)
        want = {}
        for bi in range(nb):
            for bj in range(bi, nb):
                if k % step == 0:
                    want[k] = (bi, bj)
                k += 1
        for t, ij in want.items():
            assert tri_tile(t, nb) == ij


def _slab_epilogue_emulated(rank, world, planes_mine, f10, f01, g2, normalization):
    """Host stand-in of sgpu_slab_raw / sgpu_slab_finalize with the SAME collective pattern as SlabEpilogue.run: the
    planes of all ranks are visible to every rank (all_gather instead of CUDA IPC mappings), each rank sums them over
    its share of the tiles, one all-reduce MAX of {-min, max}, each rank normalises its share."""
    n = planes_mine.shape[1]
    nb = (n + 31) // 32
    got = [torch.empty_like(planes_mine) for _ in range(world)]
    dist.all_gather(got, planes_mine)
    t0, t1 = slab_tiles(n, rank, world)
    raw = np.zeros((n, n))
    mask = np.zeros((n, n), bool)
    for t in range(t0, t1):
        bi, bj = tri_tile(t, nb)
        blk = (slice(bi * 32, min(n, bi * 32 + 32)), slice(bj * 32, min(n, bj * 32 + 32)))
        acc = sum(g.numpy()[(slice(None),) + blk].astype(np.int64) for g in got)
        v = f10 * acc[0] + f01 * acc[1]
        for k in range(3):
            v = v + np.where(acc[2 + k] != 0, g2[k] * acc[2 + k], 0.0)
        up = np.triu(np.ones((n, n), bool), 1)[blk]
        raw[blk] = np.where(up, v, 0.0)
        mask[blk] = up
    ext = torch.tensor([-min(0.0, raw[mask].min() if mask.any() else 0.0), max(0.0, raw[mask].max() if mask.any() else 0.0)],
                       dtype=torch.float64)
    dist.all_reduce(ext, op=dist.ReduceOp.MAX)
    mx = float(ext[1])
    if normalization == "ADD_MIN":
        out = raw * -1 + abs(-mx)
    elif normalization == "EXPONENTIATE":
        out = 1.0 / (np.exp(raw) + 1)
    else:
        out = raw * (1.0 / mx)
    return np.where(mask, out, 0.0), mask


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
    # no read keeps more than two loci (the host stand-in of the epilogue below knows the order-2 planes only): the
    # entries of a read at its third and later loci get fresh read ids
    locus = np.repeat(np.arange(f.n_loci), np.diff(f.row_ptr.astype(np.int64)))
    chrom = np.searchsorted(f.chr_ptr.astype(np.int64), locus, side="right") - 1
    key = chrom.astype(np.int64) << 32 | f.read_id.astype(np.int64)
    order = np.lexsort((locus, key))
    ks, ls_ = key[order], locus[order]
    new_read = np.r_[True, ks[1:] != ks[:-1]]
    new_locus = new_read | np.r_[True, ls_[1:] != ls_[:-1]]
    locus_rank = np.cumsum(new_locus) - np.maximum.accumulate(np.where(new_read, np.cumsum(new_locus), 0))
    late = locus_rank >= 2
    rid = f.read_id.copy()
    rid[order[late]] = (int(f.read_id.max()) + 1 + np.arange(int(late.sum()))).astype(np.uint32)
    f = Pileup(f.chr_ptr, f.row_ptr, f.position, rid, f.gid_base)
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
    else:
        assert lists == []
    # the peer-memory epilogue (no reduction onto one rank): shares of tiles, one scalar all-reduce
    ls, ld = po.log_probs(0.01, 0.5, 0.01, 1000, 4)
    F = ld - ls
    f10, f01 = F[1, 0], F[0, 1]
    g2 = [F[2 - d, d] - (2 - d) * f10 - d * f01 for d in range(3)]
    planes_mine = torch.from_numpy(np.stack([r.S1, r.D1, r.H[0], r.H[1], r.H[2]]).astype(np.int32))
    for norm in ("ADD_MIN", "EXPONENTIATE", "SCALE_MAX_1"):
        share, mask = _slab_epilogue_emulated(rank, world, planes_mine, f10, f01, g2, norm)
        parts = [torch.zeros(n, n, dtype=torch.float64) for _ in range(world)]
        masks = [torch.zeros(n, n, dtype=torch.bool) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(share))
        dist.all_gather(masks, torch.from_numpy(mask))
        if rank == 0:
            cover = sum(m.numpy().astype(int) for m in masks)
            ok = ok and np.array_equal(cover, np.triu(np.ones((n, n), int), 1))       # every element owned exactly once
            upper = sum(x.numpy() for x in parts)
            M = upper + upper.T
            np.fill_diagonal(M, 0.0)
            want = po.normalize(whole.raw, norm)
            # classes of order >= 3 exist in this pileup but are not part of the emulation: compare where they are absent
            h3 = (whole.class_hist.sum() - whole.class_hist[:3, :3][np.add.outer(np.arange(3), np.arange(3)) <= 2].sum()) > 0
            ok = ok and not h3 and whole.H.sum() > 0
            ok = ok and np.abs(M - want).max() <= 1e-9 * max(1e-300, np.abs(want).max())
    if rank == 0:
        q.put(bool(ok))
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
