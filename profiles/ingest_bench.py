"""Throughput of the binary-pileup ingestion (sgpu_pileup_from_bin: host record walk + H2D of the raw
bytes + unpack kernel) next to the reference's read_pileup_bin on the same file (oracle/_ref, one core).
    python profiles/ingest_bench.py [n_loci]"""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: F401  (CUDA context / pinned memory)
from secedo_b200 import api
from oracle import pyoracle as po

n_loci = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = api.Context(0)
dev = ctx.synth_pileup(8000, 0.5, 1, n_loci, theta=0.001, p_multi=0.005, p_mate=0.01, seed=5)
p = dev.download()
dev.free()
raw = np.frombuffer(p.to_bin(0), np.uint8)
pinned = torch.from_numpy(raw.copy()).pin_memory().numpy()
ident = np.arange(10000, dtype=np.uint16)
for buf, label in ((raw, "pageable"), (pinned, "pinned")):
    for _ in range(2):
        d, *_ = ctx.pileup_from_bin([buf], ident, 65535)
        d.free()
    ctx.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        d, n_cells, _, _ = ctx.pileup_from_bin([buf], ident, 65535)
        d.free()
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"sgpu_pileup_from_bin [{label} host buffer]: {raw.size / 1e6:.1f} MB, {p.n_loci} loci, {p.n_entries} entries: "
          f"{dt * 1e3:.1f} ms = {raw.size / dt / 1e9:.2f} GB/s, {p.n_loci / dt:.0f} loci/s")
d, *_ = ctx.pileup_from_bin([raw], ident, 65535)
assert d.download() == p
if po.have_ref():
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "chr1.bin")
        raw.tofile(path)
        t0 = time.perf_counter()
        r, n_cells, _ = po.ref_read_pileup(path, ident, 65535)
        dt = time.perf_counter() - t0
        assert r.n_entries == p.n_entries
        print(f"reference read_pileup_bin (page-cached file, 1 core, incl. flattening to CSR): {dt * 1e3:.1f} ms = "
              f"{raw.size / dt / 1e9:.2f} GB/s, {p.n_loci / dt:.0f} loci/s")
