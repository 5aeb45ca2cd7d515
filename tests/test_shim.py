"""The C++ drop-in shim (secedo_b200/host/secedo_b200_shim.cpp): it compiles against the reference's
own headers (source compatibility, checked where /root/reference exists) and against the stand-in
headers; on a GPU box a C++ driver calls Filter::filter + computeSimilarityMatrix through it, the way
divide_cluster does, and the outputs are checked against the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_matrix_close
from oracle import pyoracle as po
from secedo_b200.pileup import NO_POS, Pileup
from secedo_b200.synth import SynthConfig, make_pileup

SHIM = os.path.join(ROOT, "secedo_b200", "host", "secedo_b200_shim.cpp")
COMPAT = os.path.join(ROOT, "tests", "compat")
INC = os.path.join(ROOT, "include")
REF = "/root/reference"


def test_shim_compiles_against_standin_headers():
    subprocess.run(["g++", "-std=c++20", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-I" + COMPAT, "-I" + INC, SHIM],
                   check=True)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "similarity_matrix.hpp")), reason="needs /root/reference")
def test_shim_compiles_against_reference_headers():
    # the reference builds with -Wall -Wextra -Werror (CMakeLists.txt:8,47); so must the replacement TU
    # include paths of the reference's own build (CMakeLists.txt: the source root and the vendored spdlog)
    subprocess.run(["g++", "-std=c++20", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-I" + REF,
                    "-I" + os.path.join(REF, "third_party", "spdlog", "include"), "-I" + INC, SHIM], check=True)


def test_shim_bench_driver_compiles():
    """the C++ driver bench.py times the shim with (e2e_shim) builds against the stand-in headers of the reference interface"""
    subprocess.run(["g++", "-std=c++20", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-I" + COMPAT, "-I" + INC,
                    os.path.join(ROOT, "tests", "cpp", "shim_bench.cpp")], check=True)


def _write_vec(f, a):
    f.write(struct.pack("<Q", a.size))
    f.write(a.tobytes())


def _read_vec(f, dtype):
    n = struct.unpack("<Q", f.read(8))[0]
    return np.frombuffer(f.read(n * np.dtype(dtype).itemsize), dtype=dtype).copy()


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["0", "all"])
def test_shim_end_to_end(tmp_path, devices):
    """devices = "0": one GPU, with the device-resident pileup cache checked by the driver (one upload for the whole
    divide_cluster-style call sequence); "all": computeSimilarityMatrix spread over every visible GPU (sgpu_multi_*)"""
    import torch
    if devices == "all" and torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ)
    if devices == "0":
        env["SECEDO_B200_DEVICES"] = "0"
    else:
        env.pop("SECEDO_B200_DEVICES", None)
    exe = str(tmp_path / "shim_driver")
    subprocess.run(["g++", "-std=c++20", "-O2", "-I" + COMPAT, "-I" + INC, os.path.join(ROOT, "tests", "cpp", "shim_driver.cpp"),
                    SHIM, "-L" + os.path.join(ROOT, "secedo_b200"), "-lsecedo_b200",
                    "-Wl,-rpath," + os.path.join(ROOT, "secedo_b200"), "-o", exe], check=True)
    cfg = SynthConfig(n_cells=120, coverage=0.3, n_loci=1500, n_chr=3, n_clones=2, p_multi=0.2, p_mate=0.1, seed=77)
    p = make_pileup(cfg)
    members = np.arange(0, 120, dtype=np.uint32)
    id_to_pos = np.full(120, NO_POS, np.uint32)
    keep = np.r_[0:50, 60:110]                      # a sub-cluster of 100 cells
    id_to_pos[keep] = np.arange(keep.size)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        for a in (p.chr_ptr, p.row_ptr, p.position, p.read_id, p.gid_base, id_to_pos):
            _write_vec(f, a)
    L, T, eps, h, theta = 1000, 4, 0.01, 0.5, 0.01
    subprocess.run([exe, fin, fout, str(keep.size), str(L), str(T), str(eps), str(h), str(theta), "ADD_MIN"], check=True, env=env)
    with open(fout, "rb") as f:
        got = Pileup(_read_vec(f, np.uint64), _read_vec(f, np.uint64), _read_vec(f, np.uint32), _read_vec(f, np.uint32),
                     _read_vec(f, np.uint16))
        cov = _read_vec(f, np.float64)[0]
        M = _read_vec(f, np.float64).reshape(keep.size, keep.size)
        prob = _read_vec(f, np.float64)
    kl, ke, _, cov64 = po.filter_flags(p, id_to_pos, theta)
    want = p.select(kl, ke)
    assert got == want and cov == cov64
    o = po.similarity(want, keep.size, L, id_to_pos, eps, h, theta, T, "ADD_MIN")
    assert_matrix_close(M, o.M, 1e-6)
    # expectation_maximization through the shim's symbol (root cluster: identity map)
    kl, ke, _, _ = po.filter_flags(p, members, theta)
    start = np.where(np.arange(120) * 7 % 10 < 5, 0.3, 0.7)
    want_prob, _ = po.expectation_maximization(p.select(kl, ke), members, theta, start)
    assert np.abs(prob - want_prob).max() <= 1e-6
