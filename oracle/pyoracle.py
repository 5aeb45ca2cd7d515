"""TEST INFRASTRUCTURE ONLY — ctypes front end for oracle/liboracle.so (the C restatement) and
oracle/_ref/libsecedo_ref.so (the unmodified reference compiled by oracle/Makefile).

May be imported from tests/, from __graft_entry__.smoke() and from bench.py's cpu_baseline /
``--impl reference`` legs only. Nothing under secedo_b200/ imports it.

All functions take objects exposing the CSR attributes of ``secedo_b200.pileup.Pileup``
(chr_ptr, row_ptr, position, read_id, gid_base) and return numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsecedo_ref.so")
MAX_CLASS = 64
NORMALIZATIONS = {"ADD_MIN": 0, "EXPONENTIATE": 1, "SCALE_MAX_1": 2}

_u8p, _u16p, _u32p, _u64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64))
_i32p, _f64p = C.POINTER(C.c_int32), C.POINTER(C.c_double)


def build(ref: bool = True) -> None:
    """Compile the restatement and, when the reference sources are present, oracle/_ref."""
    subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)
    if ref:
        subprocess.run(["make", "-C", HERE, "ref", "-j8"], check=True, capture_output=True)


class _quiet_stdout:
    """The reference prints a ProgressBar on stdout (similarity_matrix.cpp:338-341); silence it at
    the file-descriptor level for the duration of a call."""

    def __enter__(self):
        import sys
        sys.stdout.flush()
        self._saved = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)

    def __exit__(self, *exc):
        os.dup2(self._saved, 1)
        os.close(self._null)
        os.close(self._saved)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        _oracle = C.CDLL(ORACLE_SO)
        _oracle.orc_similarity.restype = C.c_int
        _oracle.orc_filter.restype = C.c_int
        _oracle.orc_similarity_wide.restype = C.c_int
        _oracle.orc_filter_wide.restype = C.c_int
        _oracle.orc_log_probs.restype = C.c_int
        _oracle.orc_normalize.restype = C.c_int
        _oracle.orc_laplacian.restype = C.c_int
        _oracle.orc_expectation_maximization.restype = C.c_int
        _oracle.orc_is_significant.restype = C.c_int
    return _oracle


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        _ref = C.CDLL(REF_SO)
        _ref.ref_filter.restype = C.c_double
        _ref.ref_similarity.restype = C.c_double
        _ref.ref_read_pileup_text.restype = C.c_uint32
        _ref.ref_omp_max_threads.restype = C.c_int
        if hasattr(_ref, "ref_expectation_maximization"):
            _ref.ref_expectation_maximization.restype = C.c_double
    return _ref


def _csr(p):
    return (C.c_uint32(p.n_chr), _p(p.chr_ptr, _u64p), _p(p.row_ptr, _u64p), _p(p.position, _u32p),
            _p(p.read_id, _u32p), _p(p.gid_base, _u16p))


# ------------------------------------------------------------------------------- restatement
def is_significant(counts4, theta: float, cell_proportion: int = 4) -> np.ndarray:
    counts4 = np.ascontiguousarray(counts4, np.uint16).reshape(-1, 4)
    lib = oracle_lib()
    out = np.zeros(counts4.shape[0], np.uint8)
    for i in range(counts4.shape[0]):
        out[i] = lib.orc_is_significant(_p(counts4[i], _u16p), C.c_double(theta), C.c_int(cell_proportion))
    return out


def filter_flags(p, id_to_pos, theta: float, cell_proportion: int = 4):
    """Returns (keep_locus, keep_entry, avg_coverage_like_reference, avg_coverage_64bit)."""
    id_to_pos = np.ascontiguousarray(id_to_pos, np.uint32)
    kl, ke = np.zeros(p.n_loci, np.uint8), np.zeros(p.n_entries, np.uint8)
    nl, ne = C.c_uint64(), C.c_uint64()
    cov, cov64 = C.c_double(), C.c_double()
    wide = p.gid_base.dtype == np.uint32  # group ids beyond the reference's 14 bits: NO_POS is 0xFFFFFFFF there
    fn = oracle_lib().orc_filter_wide if wide else oracle_lib().orc_filter
    rc = fn(C.c_uint32(p.n_chr), _p(p.chr_ptr, _u64p), _p(p.row_ptr, _u64p),
                                 _p(p.read_id, _u32p), _p(p.gid_base, _u32p if wide else _u16p), _p(id_to_pos, _u32p),
                                 C.c_uint32(id_to_pos.size), C.c_double(theta), C.c_int(cell_proportion),
                                 _p(kl, _u8p), _p(ke, _u8p), C.byref(nl), C.byref(ne), C.byref(cov),
                                 C.byref(cov64))
    if rc:
        raise ValueError(f"orc_filter rc={rc}")
    return kl, ke, cov.value, cov64.value


def log_probs(mutation_rate, homozygous_rate, seq_error_rate, max_fragment_length, n):
    ls, ld = np.zeros((n, n)), np.zeros((n, n))
    rc = oracle_lib().orc_log_probs(C.c_double(mutation_rate), C.c_double(homozygous_rate),
                                    C.c_double(seq_error_rate), C.c_uint32(max_fragment_length),
                                    C.c_uint32(n), _p(ls, _f64p), _p(ld, _f64p))
    if rc:
        raise ValueError(f"orc_log_probs rc={rc}")
    return ls, ld


def similarity(p, num_cells, max_fragment_length, group_id_to_pos, mutation_rate, homozygous_rate,
               seq_error_rate, num_threads, normalization="ADD_MIN", instrument=True):
    """Returns a namespace with M (normalised), and when instrument: S1, D1, H[3], class_hist,
    K (per chromosome), raw (un-normalised mat_diff - mat_same)."""
    g = np.ascontiguousarray(group_id_to_pos, np.uint32)
    n = int(num_cells)
    M = np.zeros((n, n))
    r = SimpleNamespace(M=M, S1=None, D1=None, H=None, class_hist=None, K=None, raw=None)
    if instrument:
        r.S1, r.D1 = np.zeros((n, n), np.int32), np.zeros((n, n), np.int32)
        r.H = np.zeros((3, n, n), np.int32)
        r.class_hist = np.zeros((MAX_CLASS, MAX_CLASS), np.uint64)
        r.K = np.zeros(max(p.n_chr, 1), np.uint64)
        r.raw = np.zeros((n, n))
    wide = p.gid_base.dtype == np.uint32
    fn = oracle_lib().orc_similarity_wide if wide else oracle_lib().orc_similarity
    csr = _csr(p)[:5] + (_p(p.gid_base, _u32p),) if wide else _csr(p)
    rc = fn(
        *csr, C.c_uint32(n), C.c_uint32(max_fragment_length), _p(g, _u32p), C.c_uint32(g.size),
        C.c_double(mutation_rate), C.c_double(homozygous_rate), C.c_double(seq_error_rate),
        C.c_uint32(num_threads), C.c_int(NORMALIZATIONS[normalization]), _p(M, _f64p), _p(r.S1, _i32p),
        _p(r.D1, _i32p), _p(r.H, _i32p), _p(r.class_hist, _u64p), _p(r.K, _u64p), _p(r.raw, _f64p))
    if rc:
        raise ValueError(f"orc_similarity rc={rc}")
    if r.K is not None:
        r.K = r.K[:p.n_chr]
    return r


def normalize(m, normalization: str) -> np.ndarray:
    m = np.array(m, dtype=np.float64, order="C")
    rc = oracle_lib().orc_normalize(C.c_int(NORMALIZATIONS[normalization]), C.c_uint32(m.shape[0]), _p(m, _f64p))
    if rc:
        raise ValueError("invalid normalization")
    return m


def expectation_maximization(p, id_to_pos, theta: float, prob_cluster_b, max_iterations: int = 0):
    """expectation_maximization() restated (oracle/secedo_oracle.c). Returns (prob_cluster_b, iterations)."""
    prob = np.array(prob_cluster_b, dtype=np.float64, order="C")
    m = np.ascontiguousarray(id_to_pos, np.uint32)
    it = C.c_uint32(0)
    rc = oracle_lib().orc_expectation_maximization(
        C.c_uint32(len(p.chr_ptr) - 1), _p(p.chr_ptr, _u64p), _p(p.row_ptr, _u64p), _p(p.gid_base, _u16p), _p(m, _u32p),
        C.c_uint32(m.size), C.c_double(theta), C.c_uint32(prob.size), _p(prob, _f64p), C.c_uint32(max_iterations), C.byref(it))
    if rc:
        raise ValueError("group id outside prob_cluster_b / id_to_pos")
    return prob, int(it.value)


def ref_expectation_maximization(p, id_to_pos, theta: float, prob_cluster_b):
    """the compiled reference's expectation_maximization; returns (prob_cluster_b, seconds)"""
    prob = np.array(prob_cluster_b, dtype=np.float64, order="C")
    m = np.ascontiguousarray(id_to_pos, np.uint32)
    with _quiet_stdout():
        secs = ref_lib().ref_expectation_maximization(
            C.c_uint32(len(p.chr_ptr) - 1), _p(p.chr_ptr, _u64p), _p(p.row_ptr, _u64p), _p(p.position, _u32p),
            _p(p.read_id, _u32p), _p(p.gid_base, _u16p), _p(m, _u32p), C.c_uint32(m.size), C.c_double(theta),
            C.c_uint32(prob.size), _p(prob, _f64p))
    return prob, float(secs)


def laplacian(a) -> np.ndarray:
    """laplacian() of the reference (spectral_clustering.cpp:33-52), restated in oracle/secedo_oracle.c."""
    a = np.ascontiguousarray(a, np.float64)
    out = np.zeros_like(a)
    rc = oracle_lib().orc_laplacian(C.c_uint32(a.shape[0]), _p(a, _f64p), _p(out, _f64p))
    assert rc == 0
    return out


def spectral_embedding(similarity, k: int):
    """What spectral_clustering() takes from ``arma::eig_sym(eigenvalues, eigenvectors, lap)``
    (spectral_clustering.cpp:127-138): all eigenpairs of the Laplacian, ascending; the first k are returned.
    Armadillo 10.3.0 (third_party/armadillo-10.3.0) forwards eig_sym to LAPACK ``dsyev``/``dsyevd``; LAPACK is not
    vendored by the reference, so the same published routine is reached here through numpy.linalg.eigh (``dsyevd``
    of the OpenBLAS bundled with numpy). Eigenvector signs are arbitrary in both."""
    w, v = np.linalg.eigh(laplacian(similarity))
    return w[:k].copy(), np.ascontiguousarray(v[:, :k])


# --------------------------------------------------------------------- compiled reference
def ref_is_significant(counts4, theta: float, cell_proportion: int = 4) -> np.ndarray:
    counts4 = np.ascontiguousarray(counts4, np.uint16).reshape(-1, 4)
    out = np.zeros(counts4.shape[0], np.uint8)
    ref_lib().ref_is_significant(_p(counts4, _u16p), C.c_uint64(counts4.shape[0]), C.c_double(theta),
                                 C.c_int(cell_proportion), _p(out, _u8p))
    return out


def ref_filter(p, id_to_pos, theta: float, cell_proportion: int = 4, num_threads: int = 1):
    """Runs Filter::filter; returns (filtered pileup as SimpleNamespace of arrays, avg_coverage, seconds)."""
    id_to_pos = np.ascontiguousarray(id_to_pos, np.uint32)
    nl, ne, cov = C.c_uint64(), C.c_uint64(), C.c_double()
    lib = ref_lib()
    secs = lib.ref_filter(*_csr(p), _p(id_to_pos, _u32p), C.c_uint32(id_to_pos.size), C.c_double(theta),
                          C.c_int(cell_proportion), C.c_uint32(num_threads), C.byref(nl), C.byref(ne), C.byref(cov))
    out = SimpleNamespace(chr_ptr=np.zeros(p.n_chr + 1, np.uint64), row_ptr=np.zeros(nl.value + 1, np.uint64),
                          position=np.zeros(nl.value, np.uint32), read_id=np.zeros(ne.value, np.uint32),
                          gid_base=np.zeros(ne.value, np.uint16))
    lib.ref_filter_fetch(_p(out.chr_ptr, _u64p), _p(out.row_ptr, _u64p), _p(out.position, _u32p),
                         _p(out.read_id, _u32p), _p(out.gid_base, _u16p))
    return out, cov.value, secs


def ref_similarity(p, num_cells, max_fragment_length, group_id_to_pos, mutation_rate, homozygous_rate,
                   seq_error_rate, num_threads, normalization="ADD_MIN"):
    """Runs computeSimilarityMatrix; returns (M, seconds)."""
    g = np.ascontiguousarray(group_id_to_pos, np.uint32)
    n = int(num_cells)
    M = np.zeros((n, n))
    with _quiet_stdout():
      secs = ref_lib().ref_similarity(
        *_csr(p), C.c_uint32(n), C.c_uint32(max_fragment_length), _p(g, _u32p), C.c_uint32(g.size),
        C.c_double(mutation_rate), C.c_double(homozygous_rate), C.c_double(seq_error_rate),
        C.c_uint32(num_threads), normalization.encode(), _p(M, _f64p))
    return M, secs


def ref_log_probs(mutation_rate, homozygous_rate, seq_error_rate, max_fragment_length, n):
    ls, ld = np.zeros((n, n)), np.zeros((n, n))
    ref_lib().ref_log_probs(C.c_double(mutation_rate), C.c_double(homozygous_rate), C.c_double(seq_error_rate),
                            C.c_uint32(max_fragment_length), C.c_uint32(n), _p(ls, _f64p), _p(ld, _f64p))
    return ls, ld


def ref_read_pileup_text(path: str, max_coverage: int = 100):
    """Runs the reference's read_pileup on a text file; returns (arrays namespace, max_fragment_length)."""
    nl, ne = C.c_uint64(), C.c_uint64()
    lib = ref_lib()
    max_len = lib.ref_read_pileup_text(path.encode(), C.c_uint32(max_coverage), C.byref(nl), C.byref(ne))
    out = SimpleNamespace(chr_ptr=np.zeros(2, np.uint64), row_ptr=np.zeros(nl.value + 1, np.uint64),
                          position=np.zeros(nl.value, np.uint32), read_id=np.zeros(ne.value, np.uint32),
                          gid_base=np.zeros(ne.value, np.uint16))
    lib.ref_filter_fetch(_p(out.chr_ptr, _u64p), _p(out.row_ptr, _u64p), _p(out.position, _u32p),
                         _p(out.read_id, _u32p), _p(out.gid_base, _u16p))
    out.n_chr, out.n_loci, out.n_entries = 1, int(nl.value), int(ne.value)
    return out, int(max_len)


def ref_read_pileup(path: str, id_to_group, max_coverage: int = 100, positions=()):
    """The reference's read_pileup (.bin or text by suffix) with a grouping and an optional position list;
    returns (arrays namespace, n_cells, max_fragment_length)."""
    nl, ne, ml = C.c_uint64(), C.c_uint64(), C.c_uint32()
    lib = ref_lib()
    g = np.ascontiguousarray(id_to_group, np.uint16)
    pos = np.ascontiguousarray(positions, np.uint32)
    lib.ref_read_pileup.restype = C.c_uint32
    n_cells = lib.ref_read_pileup(path.encode(), _p(g, C.POINTER(C.c_uint16)), C.c_uint32(g.size), C.c_uint32(max_coverage),
                                  _p(pos, _u32p), C.c_uint64(pos.size), C.byref(nl), C.byref(ne), C.byref(ml))
    out = SimpleNamespace(chr_ptr=np.zeros(2, np.uint64), row_ptr=np.zeros(nl.value + 1, np.uint64),
                          position=np.zeros(nl.value, np.uint32), read_id=np.zeros(ne.value, np.uint32),
                          gid_base=np.zeros(ne.value, np.uint16))
    lib.ref_filter_fetch(_p(out.chr_ptr, _u64p), _p(out.row_ptr, _u64p), _p(out.position, _u32p),
                         _p(out.read_id, _u32p), _p(out.gid_base, _u16p))
    out.n_chr, out.n_loci, out.n_entries = 1, int(nl.value), int(ne.value)
    return out, int(n_cells), int(ml.value)


def ref_omp_max_threads() -> int:
    return ref_lib().ref_omp_max_threads()
