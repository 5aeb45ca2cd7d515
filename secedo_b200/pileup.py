"""Flat (CSR) pileup container — the host-side data type on the boundary.

The reference keeps a pileup as ``std::vector<std::vector<PosData>>`` (one vector per chromosome,
one ``PosData{position, read_ids[], group_ids_bases[]}`` per locus, ``sequenced_data.hpp:11-47``).
The C ABI (``include/secedo_b200.h``) takes the same information as five flat arrays so that it
can be staged to HBM with five copies:

    chr_ptr[n_chr+1]   u64  locus offsets per chromosome
    row_ptr[n_loci+1]  u64  entry offsets per locus
    position[n_loci]   u32  genomic position, strictly increasing inside a chromosome
    read_id[n_entries] u32  read (fragment) id, unique inside a chromosome
    gid_base[n_entries] u16 group id << 2 | base   (14-bit group id, ``sequenced_data.hpp:29-36``)

A pileup whose ``gid_base`` array is uint32 is WIDE: group ids beyond the reference's 14 bits (BASELINE config 5,
20 000 cells); "not in the cluster" is then ``NO_POS_WIDE`` instead of ``NO_POS`` in the id maps.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

NO_POS = 16383  # util/is_significant.hpp:11
NO_POS_WIDE = 0xFFFFFFFF
_CHAR_TO_INT = {"A": 0, "a": 0, "C": 1, "c": 1, "G": 2, "g": 2, "T": 3, "t": 3}  # util/util.hpp:17-22


@dataclass
class Pileup:
    chr_ptr: np.ndarray
    row_ptr: np.ndarray
    position: np.ndarray
    read_id: np.ndarray
    gid_base: np.ndarray

    def __post_init__(self):
        self.chr_ptr = np.ascontiguousarray(self.chr_ptr, dtype=np.uint64)
        self.row_ptr = np.ascontiguousarray(self.row_ptr, dtype=np.uint64)
        self.position = np.ascontiguousarray(self.position, dtype=np.uint32)
        self.read_id = np.ascontiguousarray(self.read_id, dtype=np.uint32)
        wide = np.asarray(self.gid_base).dtype == np.uint32
        self.gid_base = np.ascontiguousarray(self.gid_base, dtype=np.uint32 if wide else np.uint16)
        assert self.chr_ptr.ndim == 1 and self.chr_ptr.size >= 1
        assert self.row_ptr.size == self.n_loci + 1
        assert self.position.size == self.n_loci
        assert self.read_id.size == self.n_entries == self.gid_base.size

    # ------------------------------------------------------------------ sizes
    @property
    def wide(self) -> bool:
        return self.gid_base.dtype == np.uint32

    @property
    def n_chr(self) -> int:
        return int(self.chr_ptr.size - 1)

    @property
    def n_loci(self) -> int:
        return int(self.chr_ptr[-1])

    @property
    def n_entries(self) -> int:
        return int(self.row_ptr[-1]) if self.row_ptr.size else 0

    def __eq__(self, other) -> bool:
        return (
            isinstance(other, Pileup)
            and np.array_equal(self.chr_ptr, other.chr_ptr)
            and np.array_equal(self.row_ptr, other.row_ptr)
            and np.array_equal(self.position, other.position)
            and np.array_equal(self.read_id, other.read_id)
            and np.array_equal(self.gid_base, other.gid_base)
        )

    # ------------------------------------------------------------ construction
    @staticmethod
    def empty(n_chr: int = 0) -> "Pileup":
        return Pileup(np.zeros(n_chr + 1, np.uint64), np.zeros(1, np.uint64), np.zeros(0, np.uint32),
                      np.zeros(0, np.uint32), np.zeros(0, np.uint16))

    @staticmethod
    def from_pos_data(chromosomes: Sequence[Sequence[Tuple[int, Sequence[int], Sequence[int]]]]) -> "Pileup":
        """Build from ``[[(position, read_ids, group_ids_bases), ...] per chromosome]`` — the
        literal shape of the reference's ``vector<vector<PosData>>``."""
        chr_ptr, row_ptr, pos, rid, gb = [0], [0], [], [], []
        for chrom in chromosomes:
            for position, read_ids, gids_bases in chrom:
                assert len(read_ids) == len(gids_bases)
                pos.append(position)
                rid.extend(read_ids)
                gb.extend(gids_bases)
                row_ptr.append(len(rid))
            chr_ptr.append(len(pos))
        return Pileup(np.array(chr_ptr, np.uint64), np.array(row_ptr, np.uint64), np.array(pos, np.uint32),
                      np.array(rid, np.uint32), np.array(gb, np.uint16))

    def to_pos_data(self) -> List[List[Tuple[int, List[int], List[int]]]]:
        out = []
        for c in range(self.n_chr):
            chrom = []
            for l in range(int(self.chr_ptr[c]), int(self.chr_ptr[c + 1])):
                a, b = int(self.row_ptr[l]), int(self.row_ptr[l + 1])
                chrom.append((int(self.position[l]), self.read_id[a:b].tolist(), self.gid_base[a:b].tolist()))
            out.append(chrom)
        return out

    @staticmethod
    def from_text(path: str, max_coverage: int = 100) -> Tuple["Pileup", int]:
        """Parse one chromosome in the reference's textual ``.pileup`` format (tab separated:
        chromosome, position, coverage, bases, cell ids, read names; read names are mapped to
        dense ids in order of first appearance and loci with more than ``max_coverage`` reads are
        skipped — ``util/pileup_reader.cpp:12-137``). Returns the pileup and the longest fragment
        (last - first position of a read id), which the reference uses as max_fragment_length."""
        row_ptr, pos, rid, gb = [0], [], [], []
        names: dict = {}
        first_last: dict = {}
        with open(path) as f:
            for line in f:
                if not line.strip():
                    continue
                cols = line.rstrip("\n").split("\t")
                position = int(cols[1])
                bases = cols[3]
                cells = [int(x) for x in cols[4].split(",")]
                if len(cells) > max_coverage:
                    continue
                reads = [x for x in cols[5].split(",")]
                for j, b in enumerate(bases):
                    r = names.setdefault(reads[j], len(names))
                    rid.append(r)
                    gb.append((cells[j] << 2) | _CHAR_TO_INT[b])
                    fl = first_last.setdefault(r, [position, position])
                    fl[1] = position
                pos.append(position)
                row_ptr.append(len(rid))
        max_len = max((b - a for a, b in first_last.values()), default=0)
        return (Pileup(np.array([0, len(pos)], np.uint64), np.array(row_ptr, np.uint64), np.array(pos, np.uint32),
                       np.array(rid, np.uint32), np.array(gb, np.uint16)), max_len)

    # --------------------------------------------------------------- slicing
    def select(self, keep_locus: np.ndarray, keep_entry: np.ndarray) -> "Pileup":
        """Compact by 0/1 flags, preserving order (the shape of ``Filter::filter``'s output)."""
        keep_locus = np.asarray(keep_locus, dtype=bool)
        keep_entry = np.asarray(keep_entry, dtype=bool)
        lens = np.diff(self.row_ptr.astype(np.int64))
        locus_of_entry = np.repeat(np.arange(self.n_loci), lens)
        keep_entry = keep_entry & keep_locus[locus_of_entry]
        kept_per_locus = np.bincount(locus_of_entry[keep_entry], minlength=self.n_loci)[keep_locus]
        row_ptr = np.concatenate([[0], np.cumsum(kept_per_locus)]).astype(np.uint64)
        cum_loci = np.concatenate([[0], np.cumsum(keep_locus)]).astype(np.uint64)
        chr_ptr = cum_loci[self.chr_ptr.astype(np.int64)]
        return Pileup(chr_ptr, row_ptr, self.position[keep_locus], self.read_id[keep_entry], self.gid_base[keep_entry])

    def to_bin(self, chrom: int) -> bytes:
        """One chromosome in the reference's binary pileup format (written by pileup.cpp:327-332, read by
        util/pileup_reader.cpp:165-179): per locus ``u32 position, u16 coverage, u32 read_id[coverage],
        u16 (cell_id << 2 | base)[coverage]``."""
        out = bytearray()
        for l in range(int(self.chr_ptr[chrom]), int(self.chr_ptr[chrom + 1])):
            a, b = int(self.row_ptr[l]), int(self.row_ptr[l + 1])
            assert b - a < 65536, "coverage is a uint16 in the binary format"
            out += np.uint32(self.position[l]).tobytes() + np.uint16(b - a).tobytes()
            out += self.read_id[a:b].astype("<u4").tobytes() + self.gid_base[a:b].astype("<u2").tobytes()
        return bytes(out)

    @staticmethod
    def from_bin(files: Sequence[bytes], id_to_group: Sequence[int], max_coverage: int = 100,
                 positions: Sequence[Sequence[int]] = None) -> Tuple["Pileup", int, int]:
        """Host-side restatement of ``read_pileup_bin`` (util/pileup_reader.cpp:139-257) over one buffer per
        chromosome, used to check ``sgpu_pileup_from_bin``: returns (pileup, n_cells, n_groups)."""
        g = np.asarray(id_to_group, np.uint16)
        chr_ptr, row_ptr, pos, rid, gb = [0], [0], [], [], []
        max_cell = max_group = 0
        for c, buf in enumerate(files):
            buf = bytes(buf)
            plist = None if positions is None or len(positions[c]) == 0 else list(positions[c])
            pidx, off = 0, 0
            while off + 6 <= len(buf):
                p = int(np.frombuffer(buf, "<u4", 1, off)[0])
                cov = int(np.frombuffer(buf, "<u2", 1, off + 4)[0])
                ids = np.frombuffer(buf, "<u4", cov, off + 6)
                cb = np.frombuffer(buf, "<u2", cov, off + 6 + 4 * cov)
                off += 6 + 6 * cov
                if cov > max_coverage:
                    continue
                if plist is not None:
                    while pidx < len(plist) and plist[pidx] < p:
                        pidx += 1
                    if pidx == len(plist):
                        break
                    if plist[pidx] > p:
                        continue
                cells = cb >> 2
                if cells.size and int(cells.max()) >= g.size:
                    raise ValueError("Cell id is too large")
                groups = g[cells]
                if cells.size:
                    max_cell, max_group = max(max_cell, int(cells.max())), max(max_group, int(groups.max()))
                pos.append(p)
                rid.append(ids)
                gb.append((groups.astype(np.uint16) << 2) | (cb & 3))
                row_ptr.append(row_ptr[-1] + cov)
            chr_ptr.append(len(pos))
        cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)  # noqa: E731
        return (Pileup(np.array(chr_ptr, np.uint64), np.array(row_ptr, np.uint64), np.array(pos, np.uint32),
                       cat(rid, np.uint32), cat(gb, np.uint16)), max_cell + 1, max_group + 1)

    def loci_range(self, chrom: int, lo: int, hi: int) -> "Pileup":
        """Loci [lo, hi) of one chromosome as a single-chromosome pileup."""
        base = int(self.chr_ptr[chrom])
        lo, hi = base + lo, min(base + hi, int(self.chr_ptr[chrom + 1]))
        a, b = int(self.row_ptr[lo]), int(self.row_ptr[hi])
        return Pileup(np.array([0, hi - lo], np.uint64), self.row_ptr[lo:hi + 1] - self.row_ptr[lo],
                      self.position[lo:hi], self.read_id[a:b], self.gid_base[a:b])

    @staticmethod
    def concat(parts: Sequence["Pileup"]) -> "Pileup":
        """Concatenate pileups chromosome-wise (chromosomes of later parts follow earlier ones)."""
        chr_ptr, row_ptr = [np.zeros(1, np.uint64)], [np.zeros(1, np.uint64)]
        lo = eo = 0
        for p in parts:
            chr_ptr.append(p.chr_ptr[1:] + np.uint64(lo))
            row_ptr.append(p.row_ptr[1:] + np.uint64(eo))
            lo += p.n_loci
            eo += p.n_entries
        return Pileup(np.concatenate(chr_ptr), np.concatenate(row_ptr), np.concatenate([p.position for p in parts]),
                      np.concatenate([p.read_id for p in parts]), np.concatenate([p.gid_base for p in parts]))
