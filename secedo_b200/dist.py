"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).

The path shards by chromosome (independent locus ranges: the reference resets all read state at
every chromosome boundary, similarity_matrix.cpp:407-408, so no halo is needed). Each rank filters
and accumulates its chromosomes into its own integer count planes; the planes are summed with ONE
reduction per buffer (int32 planes, and — only if a read pair overlapped at >= 4 loci — the fp64
spill plane), and rank ``dst`` runs the log-likelihood epilogue. Integer sums are order independent,
so the N-GPU count matrices equal the 1-GPU ones bit for bit.

Everything here is backend agnostic (the CPU tests run it with gloo and world_size 2); only
:func:`counts_tensors` touches device pointers.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def partition_chromosomes(weights: Sequence[float], world_size: int) -> List[List[int]]:
    """Longest-processing-time assignment of chromosomes to ranks, balanced by ``weights`` (loci, or
    sum of squared coverage for the scatter path). Deterministic; every rank computes the same."""
    order = sorted(range(len(weights)), key=lambda i: (-weights[i], i))
    loads = [0.0] * world_size
    parts: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        parts[r].append(i)
        loads[r] += weights[i]
    return [sorted(p) for p in parts]


def agree_layout(planes_used: int, has_spill: bool, device=None, group=None) -> tuple:
    """Ranks may have seen different overlap classes; the reduction needs identical buffers."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return planes_used, has_spill
    t = torch.tensor([planes_used, int(has_spill)], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t[0]), bool(t[1])


def reduce_buffers(tensors: Sequence[torch.Tensor], dst: int = 0, group=None) -> None:
    """Element-wise SUM of every tensor onto rank ``dst`` (one collective per buffer)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in tensors:
        if t.numel():
            dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM, group=group)


class _CudaArray:
    """Minimal __cuda_array_interface__ carrier so torch can alias a raw device pointer."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def tensor_from_ptr(ptr: int, n: int, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
    typestr = {torch.int32: "<i4", torch.float64: "<f8", torch.int64: "<i8"}[dtype]
    if n == 0 or not ptr:
        return torch.empty(0, dtype=dtype, device=device)
    return torch.as_tensor(_CudaArray(ptr, n, typestr), device=device)


def counts_tensors(counts, device: torch.device) -> List[torch.Tensor]:
    """Torch views (no copy) of the device buffers of a :class:`secedo_b200.api.Counts`."""
    i32, n_i32, f64, n_f64, hist, n_hist = counts.buffers()
    return [tensor_from_ptr(i32, n_i32, torch.int32, device), tensor_from_ptr(f64, n_f64, torch.float64, device),
            tensor_from_ptr(hist, n_hist, torch.int64, device)]


def exchange_sparse(idx: torch.Tensor, val: torch.Tensor, dst: int = 0, group=None, lens: Optional[Sequence[int]] = None):
    """Variable-length (index, value) lists of all ranks -> rank ``dst``: returns, on ``dst``, the lists of the OTHER
    ranks as [(idx, val), ...] (int32 tensors; idx carries the bits of a uint32) and an empty list elsewhere. One
    all_gather of the lengths (unless the caller already knows them), one gather of the lists padded to the longest."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if lens is None:
        n = torch.tensor([idx.numel()], dtype=torch.int64, device=idx.device)
        got_n = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(got_n, n, group=group)
        lens = [int(x.item()) for x in got_n]
    longest = max(lens)
    if longest == 0:
        return []
    mine = torch.zeros(2 * longest, dtype=torch.int32, device=idx.device)
    mine[: idx.numel()] = idx
    mine[longest: longest + val.numel()] = val
    if rank == dst:
        got = [torch.empty_like(mine) for _ in range(world)]
        dist.gather(mine, got, dst=dst, group=group)
        return [(g[: lens[r]], g[longest: longest + lens[r]]) for r, g in enumerate(got) if r != dst and lens[r]]
    dist.gather(mine, None, dst=dst, group=group)
    return []


MAX_WORLD_SPARSE = 4  # largest world size at which the sparse route is tried (see reduce_counts)


def sparse_pays(nnz_per_rank: Sequence[int], dst: int, n_sparse_planes: int, num_cells: int) -> bool:
    """The lists of the other ranks (8 bytes per non-zero into ``dst``) against a dense reduction of the packed planes."""
    others = sum(nnz_per_rank) - nnz_per_rank[dst]
    dense = 4 * n_sparse_planes * num_cells * (num_cells - 1) // 2
    return n_sparse_planes > 0 and 8 * others < dense // 2


def reduce_counts(counts, device: torch.device, dst: int = 0, group=None) -> None:
    """Agree on the buffer layout, then sum the count planes of all ranks onto ``dst``. Only upper triangles travel.
    S and D are packed into one contiguous buffer on every rank (``sgpu_counts_pack_range``), summed with ONE reduction
    and unpacked on ``dst``. The planes of the read pairs that overlap at >= 2 loci hold few non-zeros on real
    pileups: where that pays (:func:`sparse_pays`) every rank sends their non-zeros as a list
    (``sgpu_counts_sparse_pack``) and ``dst`` adds the lists into its own planes - 8 bytes per non-zero instead of
    4 bytes per cell pair and plane."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    world = dist.get_world_size(group)
    i32, n_i32, f64, n_f64, hist, n_hist = counts.buffers()
    nc = counts.num_cells
    nn = nc * nc
    local_planes = n_i32 // nn if nn else 2
    # non-zeros of this rank's second / third order planes (cheap: two passes over planes that are mostly zero)
    ip = vp = None
    nnz = 0
    force = os.environ.get("SECEDO_B200_SPARSE_REDUCE")  # tests: "1" / "0" force one route on every rank
    # The lists of all other ranks converge on ``dst`` while a dense NCCL reduce costs the same at any world size:
    # measured at 8 000 cells, 5.6 M non-zeros per rank - 1.7 ms (lists) vs 1.9 ms (dense) at 2 ranks, no gain at 8
    # (profiles/r1_reduce_probe_2gpu.txt). Beyond 4 ranks the lists are not even built.
    try_sparse = (world <= MAX_WORLD_SPARSE or force == "1") and force != "0"
    if try_sparse and local_planes > 2 and 7 * nn <= 0xFFFFFFFF:
        ip, vp, nnz = counts.sparse_pack(2)
    # ONE small collective: planes in use, spill plane, list length of every rank
    info = torch.tensor([local_planes, int(n_f64 > 0), nnz], dtype=torch.int64, device=device)
    infos = [torch.zeros_like(info) for _ in range(world)]
    dist.all_gather(infos, info, group=group)
    infos = [[int(v) for v in t.tolist()] for t in infos]
    planes, spill = max(i[0] for i in infos), any(i[1] for i in infos)
    all_nnz = [i[2] for i in infos]
    counts.set_layout(planes, spill)
    use_sparse = try_sparse and planes > 2 and 7 * nn <= 0xFFFFFFFF and sparse_pays(all_nnz, dst, planes - 2, nc)
    if force == "1" and planes > 2 and 7 * nn <= 0xFFFFFFFF:
        use_sparse = True
    idx_t = val_t = None
    if use_sparse:
        idx_t, val_t = tensor_from_ptr(ip, nnz, torch.int32, device), tensor_from_ptr(vp, nnz, torch.int32, device)
    dense_planes = 2 if use_sparse else planes
    ptr, n = counts.pack_range(0, dense_planes)
    counts.ctx.synchronize()  # the library's stream need not be the one the collective is ordered on
    _, _, f64, n_f64, hist, n_hist = counts.buffers()
    reduce_buffers([tensor_from_ptr(ptr, n, torch.int32, device), tensor_from_ptr(f64, n_f64, torch.float64, device),
                    tensor_from_ptr(hist, n_hist, torch.int64, device)], dst, group)
    lists = exchange_sparse(idx_t, val_t, dst, group, lens=all_nnz) if use_sparse else []
    if dist.get_rank(group) == dst:
        if device is not None and torch.device(device).type == "cuda":
            torch.cuda.current_stream(device).synchronize()  # the sums have arrived before they are unpacked
        counts.unpack_range(0, dense_planes)
        for li, lv in lists:
            counts.sparse_add(2, li.data_ptr(), lv.data_ptr(), li.numel())
        if lists:
            counts.ctx.synchronize()  # the gathered lists may be released
