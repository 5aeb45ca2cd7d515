"""N-GPU == 1-GPU parity over NCCL (needs >= 2 visible GPUs; skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.gpu
def test_two_gpus_match_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(ROOT, "tests", "multi_gpu_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("OK") == 9


@pytest.mark.gpu
def test_single_process_multi_gpu_similarity():
    """sgpu_multi_similarity: all visible GPUs behind one call of the C ABI (one host thread per GPU, pieces of
    chromosomes with halos, peer-memory epilogue, no NCCL). Equal to one GPU: bit for bit without read pairs of order >= 4,
    to rounding with them; and to the oracle within 1e-6 * max|M|. On a single-GPU box it runs with one piece."""
    import numpy as np
    from conftest import assert_matrix_close
    from oracle import pyoracle as po
    from secedo_b200 import api
    from secedo_b200.synth import SynthConfig, make_pileup
    mc = api.MultiContext()
    assert mc.size >= 1
    ctx = api.Context(0)
    for p_multi, exact in ((0.02, True), (0.4, False)):
        cfg = SynthConfig(n_cells=500, coverage=0.3, n_loci=900, n_chr=3, n_clones=3, p_multi=p_multi, p_mate=0.05, seed=8)
        p = make_pileup(cfg)
        ident = np.arange(cfg.n_cells, dtype=np.uint32)
        f, _ = api.Filter(0.01, 4, ctx).filter(p, ident, "", 1)
        for path in ("gemm", "scatter"):
            for norm in ("ADD_MIN", "EXPONENTIATE", "SCALE_MAX_1"):
                M, st = mc.similarity(f, cfg.n_cells, 1000, ident, 0.01, 0.5, 0.01, 4, norm, path, return_stats=True)
                one = api.compute_similarity_matrix(f, cfg.n_cells, 1000, ident, 0.01, 0.5, 0.01, 4, "", norm, ctx=ctx, path=path)
                if exact:
                    assert np.array_equal(M, one, equal_nan=True), (path, norm)
                else:
                    assert np.nanmax(np.abs(M - one)) <= 1e-12 * np.nanmax(np.abs(one)), (path, norm)
                assert st["path_used"] == path
            o = po.similarity(f, cfg.n_cells, 1000, ident, 0.01, 0.5, 0.01, 4, "ADD_MIN")
            assert_matrix_close(mc.similarity(f, cfg.n_cells, 1000, ident, 0.01, 0.5, 0.01, 4, "ADD_MIN", path), o.M, 1e-6)
    # an empty pileup and more GPUs than loci
    from secedo_b200.pileup import Pileup
    assert not mc.similarity(Pileup.empty(2), 7, 1000, np.arange(7, dtype=np.uint32), 0.01, 0.5, 0.01, 1).any()
    mc.close()
    ctx.close()
