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


# ------------------------------------------------------------------------------------------------------------------
# Multi-GPU epilogue over peer memory: nothing is reduced onto one rank. Every rank owns an equal share of the 32 x 32
# tiles of the upper triangle, reads the count planes of ALL ranks for its tiles through NVLink peer mappings (CUDA IPC
# between the processes of one node), adds them and applies the log-likelihood transform in the same kernel
# (sgpu_slab_raw), exchanges ONE scalar pair (all-reduce MAX of {-min, max}; also the barrier after which nobody reads a
# peer any more) and writes its share of the normalised matrix (sgpu_slab_finalize) — into its own HBM, or straight
# into a host matrix shared by the ranks, each over its own PCIe link.
# ------------------------------------------------------------------------------------------------------------------
def tri_tile_count(num_cells: int) -> int:
    nb = (num_cells + 31) // 32
    return nb * (nb + 1) // 2


def slab_tiles(num_cells: int, slab: int, n_slabs: int) -> tuple:
    """[first, one past last) tile of share ``slab`` (tiles numbered row by row over bj >= bi): equal counts, so equal
    work and equal NVLink traffic for every rank. Mirrors sgpu_slab_raw."""
    t = tri_tile_count(num_cells)
    return t * slab // n_slabs, t * (slab + 1) // n_slabs


def tri_tile(t: int, nb: int) -> tuple:
    """tile number -> (bi, bj) of the upper triangle of nb x nb tiles"""
    bi = 0
    before = lambda r: r * nb - r * (r - 1) // 2  # noqa: E731
    lo, hi = 0, nb
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if before(mid) <= t:
            lo = mid
        else:
            hi = mid
    bi = lo
    return bi, bi + t - before(bi)


def _as_i64(v: int) -> int:
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


class SharedHostMatrix:
    """n x n fp64 matrix in POSIX shared memory, mapped by every rank of the node and page-locked + mapped into each
    rank's GPU, so that every GPU writes its share of the result straight into the host memory rank ``dst`` reads."""

    def __init__(self, ctx, num_cells: int, group=None, dst: int = 0):
        from multiprocessing import shared_memory
        self.ctx, self.n = ctx, int(num_cells)
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        nbytes = max(8, self.n * self.n * 8)
        name = [None]
        if rank == dst:
            self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
            name[0] = self.shm.name
        if world > 1:
            dist.broadcast_object_list(name, src=dst, group=group)
        if rank != dst:
            self.shm = shared_memory.SharedMemory(name=name[0])
            try:  # only the creating rank unlinks the segment (Python < 3.13 would try it from every process at exit)
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:  # noqa: BLE001
                pass
        self.owner = rank == dst
        self.array = np.ndarray((self.n, self.n), dtype=np.float64, buffer=self.shm.buf)
        self.host_ptr = self.array.ctypes.data
        self.dev_ptr = ctx.host_register(self.host_ptr, nbytes)

    def close(self):
        if getattr(self, "shm", None) is None:
            return
        try:
            self.ctx.host_unregister(self.host_ptr)
        except Exception:  # noqa: BLE001
            pass
        self.array = None
        self.shm.close()
        if self.owner:
            self.shm.unlink()
        self.shm = None


class SlabEpilogue:
    """The peer-memory epilogue for one :class:`secedo_b200.api.Counts` per rank. Construction exchanges the CUDA IPC
    handles of the planes once (collective); :meth:`run` is then two tiny collectives and two kernels per matrix."""

    def __init__(self, counts, device: torch.device, group=None):
        self.counts, self.device, self.group = counts, device, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._opened = []
        self.peers = self._exchange(counts.buffers()[0], spill=False)
        self.peer_spill = None  # exchanged when a spill plane first appears on some rank

    def _exchange(self, own: int, spill: bool):
        """device pointers, valid on this rank's GPU, of the planes (or spill planes) of all ranks (collective)"""
        if self.world == 1:
            return [own]
        h = torch.frombuffer(bytearray(self.counts.ipc_handle(spill)), dtype=torch.uint8).to(self.device)
        got = [torch.empty_like(h) for _ in range(self.world)]
        dist.all_gather(got, h, group=self.group)
        ptrs = []
        for r, t in enumerate(got):
            if r == self.rank:
                ptrs.append(own)
            else:
                p = self.counts.ctx.ipc_open(bytes(t.cpu().numpy().tobytes()))
                self._opened.append(p)
                ptrs.append(p)
        return ptrs

    def agree(self):
        """planes in use / spill plane of all ranks -> common layout; also the barrier 'everybody has finished
        accumulating' (the collective is ordered behind the accumulate kernels on the stream). Returns (planes, spill)."""
        _, n_i32, _, n_f64, _, _ = self.counts.buffers()
        nn = self.counts.num_cells ** 2
        planes, spill = agree_layout(n_i32 // nn if nn else 2, n_f64 > 0, self.device, self.group)
        self.counts.set_layout(planes, spill)
        if spill and self.peer_spill is None:  # every rank has a spill plane now (zero where it saw no such pair)
            self.peer_spill = self._exchange(self.counts.buffers()[2], spill=True)
        return planes, spill

    def checksums(self):
        """(sum over ranks of the checksums of their own planes, sum over ranks of the checksums of the peer-summed
        shares): equal iff every rank sees, through its peer mappings, exactly what the others accumulated."""
        self.counts.ctx.synchronize()
        self.agree()  # the same planes on every rank (the ones a rank never touched are zero), and everybody has accumulated
        own = _as_i64(self.counts.checksum())
        summed = _as_i64(self.counts.checksum(self.peers, self.rank, self.world))
        t = torch.tensor([own, summed], dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        return int(t[0].item()) & ((1 << 64) - 1), int(t[1].item()) & ((1 << 64) - 1)

    def run(self, max_fragment_length, mutation_rate, homozygous_rate, seq_error_rate, normalization, out_ptr=None,
            same_stream: bool = False):
        """Returns the pointer this rank's share was written to (``out_ptr``, or a device matrix owned by the counts
        object of which only this rank's tiles and their mirror images are valid). ``same_stream``: the library's
        stream is torch's current stream, so nothing has to be synchronised on the host."""
        c = self.counts
        if not same_stream:
            c.ctx.synchronize()
        planes, spill = self.agree()
        ext = c.slab_raw(self.peers, self.rank, self.world, max_fragment_length, mutation_rate, homozygous_rate, seq_error_rate,
                         peer_spill=self.peer_spill if spill else None)
        if self.world > 1:
            if not same_stream:
                c.ctx.synchronize()
            dist.all_reduce(tensor_from_ptr(ext, 2, torch.float64, self.device), op=dist.ReduceOp.MAX, group=self.group)
            if not same_stream:
                torch.cuda.current_stream(self.device).synchronize()
        return c.slab_finalize(normalization, out_ptr)

    def close(self):
        for p in self._opened:
            try:
                self.counts.ctx.ipc_close(p)
            except Exception:  # noqa: BLE001
                pass
        self._opened = []


# ------------------------------------------------------------------------------------------------------------------
# Pieces of chromosomes: contiguous locus ranges cut INSIDE chromosomes, balanced by weight, each with a
# read-only halo of max_fragment_length bp on both sides (sgpu_counts_accumulate_range). Chromosome lengths are far too
# uneven (249 Mbp .. 47 Mbp) for whole chromosomes to balance 8 GPUs, and a pileup of one chromosome could not use
# more than one GPU at all.
# ------------------------------------------------------------------------------------------------------------------
NO_TAIL = 0xFFFFFFFF


def plan_pieces(positions: Sequence[np.ndarray], n_pieces: int, max_fragment_length: int,
                weights: Optional[Sequence[np.ndarray]] = None) -> List[List[dict]]:
    """Cut the loci of all chromosomes (``positions[c]``: ascending positions of chromosome c) into ``n_pieces`` runs of
    (nearly) equal total weight (``weights[c][l]``, default 1 per locus; use the squared coverage for the scatter path).
    Returns, per piece, a list of dicts, one per chromosome the piece touches:
        chrom, own_pos_begin, own_pos_end   owned loci = positions in [own_pos_begin, own_pos_end)
        own_lo, own_hi                      the same as locus indices of the chromosome
        lo, hi                              loci the piece has to hold: the owned ones plus the halo
    The owned ranges tile every chromosome; a piece owns at most one range per chromosome."""
    L = int(max_fragment_length)
    sizes = [int(len(p)) for p in positions]
    w = [np.ones(n, np.float64) if weights is None else np.asarray(weights[c], np.float64) for c, n in enumerate(sizes)]
    total = float(sum(x.sum() for x in w))
    cuts = [total * k / n_pieces for k in range(n_pieces + 1)]
    pieces: List[List[dict]] = [[] for _ in range(n_pieces)]
    before = 0.0
    for c, pos in enumerate(positions):
        pos = np.asarray(pos, np.int64)
        n = sizes[c]
        if n == 0:
            continue
        cum = before + np.concatenate([[0.0], np.cumsum(w[c])])  # weight in front of locus l
        for k in range(n_pieces):
            # the loci whose starting weight lies in [cuts[k], cuts[k + 1]); the last piece takes everything left
            a = int(np.searchsorted(cum[:-1], cuts[k], side="left"))
            b = n if k == n_pieces - 1 else int(np.searchsorted(cum[:-1], cuts[k + 1], side="left"))
            if b <= a:
                continue
            pb = int(pos[a]) if a > 0 else 0
            pe = int(pos[b]) if b < n else NO_TAIL
            lo = int(np.searchsorted(pos, pos[a] - L, side="right"))       # positions > first owned - L
            hi = int(np.searchsorted(pos, pos[b - 1] + L, side="left"))     # positions < last owned + L
            pieces[k].append(dict(chrom=c, own_pos_begin=pb, own_pos_end=pe, own_lo=a, own_hi=b, lo=min(lo, a), hi=max(hi, b)))
        before = float(cum[-1])
    return pieces
