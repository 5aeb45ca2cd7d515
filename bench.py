#!/usr/bin/env python
"""Benchmark of the similarity-matrix hot path (BASELINE.json metric: similarity-matrix build time &
loci/s at 8K cells, 0.5x, on 1/2/4/8 B200, next to the reference's OpenMP CPU path).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched with torchrun)
    python bench.py --impl reference --steps K --warmup W    (the reference's own CPU path)

A step = one pass of the hot path over one batch of synthetic pileup per GPU:
    Filter::filter -> read linking / mate rule / cutoff -> first-order counts (int8 tcgen05 GEMM)
    -> multi-locus correction -> [NCCL reduce of the count planes] -> log-likelihood epilogue.
Weak scaling: every rank owns its own chromosomes (its own seed); rank 0 produces the N x N matrix.
`value` times the steps with the batch already resident in HBM; `e2e` times the same call sequence
with HOST (pinned) buffers: the H2D copy of the whole batch and the D2H copy of the matrix are inside
the timed region. The CPU baseline / reference arm run the UNMODIFIED reference compiled into
oracle/_ref (or, if that .so is absent, the oracle port) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ---- workload: BASELINE.json configs[2] (headline; one batch of it fits one GPU) ----------------------
WORKLOAD = dict(
    name="cfg3: 8000 cells x 0.5x, batch of whole-genome-style pileup (4 chromosomes x %d pre-filter loci per GPU and step)",
    n_cells=int(os.environ.get("SECEDO_BENCH_CELLS", 8000)),
    coverage=float(os.environ.get("SECEDO_BENCH_COVERAGE", 0.5)),
    n_chr=4,
    loci_per_chr=int(os.environ.get("SECEDO_BENCH_LOCI_PER_CHR", 32768)),
    n_clones=2, frac_somatic=0.5, frac_germline=0.1, spacing=400,
    p_multi=float(os.environ.get("SECEDO_BENCH_P_MULTI", 0.005)),
    p_mate=float(os.environ.get("SECEDO_BENCH_P_MATE", 0.01)), p_mate_mismatch=0.2,
    # flags_sim of the reference: h = 0.15, theta = 0.001, eps = 0.01; with theta = 0.01 the reference's
    # filter accepts pure noise at this pooled coverage
    theta=0.001, eps=0.01, h=0.15, L=1000, num_threads=8, normalization="ADD_MIN",
)
CPU_SAMPLE_LOCI = int(os.environ.get("SECEDO_BENCH_CPU_LOCI", 28))  # pre-filter loci of the CPU sample


def clocks_sampler(stop, out, device_index):
    """nvidia-smi sampled during the timed region (B200_PROFILING.md)."""
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", f"--id={device_index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                              "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
        return
    def reader():
        for line in p.stdout:
            out.append(line.strip())
    t = threading.Thread(target=reader, daemon=True)
    t.start()
    stop.wait()
    p.terminate()


def summarize_clocks(lines):
    sm, mx, reasons = [], 0, set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in lines:
        f = [x.strip() for x in ln.split(",")]
        if len(f) < 7:
            continue
        try:
            sm.append(float(f[0]))
            mx = max(mx, float(f[1]))
        except ValueError:
            continue
        for name, v in zip(names, f[3:7]):
            if v.lower().startswith("active"):
                reasons.add(name)
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
            "samples": len(sm)}


# ---------------------------------------------------------------------------------------- reference arm
def cpu_sample(seed=77):
    """A bounded sample of the workload, generated on the host (same model as the device generator)."""
    from secedo_b200.synth import SynthConfig, make_pileup
    w = WORKLOAD
    cfg = SynthConfig(n_cells=w["n_cells"], coverage=w["coverage"], n_loci=CPU_SAMPLE_LOCI, n_chr=1,
                      n_clones=w["n_clones"], frac_somatic=w["frac_somatic"], frac_germline=w["frac_germline"],
                      theta=w["theta"], spacing=w["spacing"], max_fragment_length=w["L"], p_multi=w["p_multi"],
                      p_mate=w["p_mate"], p_mate_mismatch=w["p_mate_mismatch"], seed=seed)
    return make_pileup(cfg)


def run_reference_step(p, threads, want_matrix=False):
    """Filter::filter + computeSimilarityMatrix of the reference on pileup p. Returns (seconds,
    significant loci, kind[, matrix])."""
    from oracle import pyoracle as po
    from secedo_b200.pileup import Pileup
    w = WORKLOAD
    ident = np.arange(w["n_cells"], dtype=np.uint32)
    if po.have_ref():
        t0 = time.perf_counter()
        rf, _, _ = po.ref_filter(p, ident, w["theta"], 4, threads)
        f = Pileup(rf.chr_ptr, rf.row_ptr, rf.position, rf.read_id, rf.gid_base)
        M, _ = po.ref_similarity(f, w["n_cells"], w["L"], ident, w["eps"], w["h"], w["theta"], threads, w["normalization"])
        secs = time.perf_counter() - t0
        return (secs, f.n_loci, "reference", M) if want_matrix else (secs, f.n_loci, "reference")
    t0 = time.perf_counter()
    kl, ke, _, _ = po.filter_flags(p, ident, w["theta"])
    f = p.select(kl, ke)
    o = po.similarity(f, w["n_cells"], w["L"], ident, w["eps"], w["h"], w["theta"], threads, w["normalization"],
                      instrument=False)
    secs = time.perf_counter() - t0
    return (secs, f.n_loci, "port", o.M) if want_matrix else (secs, f.n_loci, "port")


def reference_threads():
    """all host cores the process may use (torchrun exports OMP_NUM_THREADS=1; the reference takes its
    thread count as an argument, so that variable does not limit it)"""
    from oracle import pyoracle as po
    if po.have_ref():
        try:
            return max(1, len(os.sched_getaffinity(0)))
        except AttributeError:
            return max(1, os.cpu_count() or 1)
    return 1


def main_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    threads = reference_threads()
    p = cpu_sample()
    times, loci, kind = [], 0, "reference"
    for i in range(args.warmup + args.steps):
        s, loci, kind = run_reference_step(p, threads)
        if i >= args.warmup:
            times.append(s)
    total = sum(times)
    value = loci * len(times) / total
    sample = (f"{CPU_SAMPLE_LOCI} pre-filter loci ({loci} significant, {p.n_entries} pileup entries) of the workload per "
              f"step; Filter::filter + computeSimilarityMatrix, num_threads={threads}")
    line = {
        "impl": "reference", "metric": "similarity-matrix significant loci/s (8K cells, 0.5x)", "value": value,
        "unit": "loci/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 log-likelihoods over integer read-pair counts", "data": "synthetic",
        "config": {"workload": WORKLOAD["name"] % WORKLOAD["loci_per_chr"], "sample": sample,
                   "n_cells": WORKLOAD["n_cells"], "coverage": WORKLOAD["coverage"]},
        "cpu_baseline": {"value": value, "unit": "loci/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "loci/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ ours
def int8_peak_tops(torch, device):
    """Dense int8 tensor-core rate of cuBLASLt on this GPU (8192^3), the denominator the survey asks
    to measure because MEASURED_PEAKS.json only holds bf16. Returns (TOP/s, how)."""
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16 = float(peaks.get("bf16_tflops", 1590.0))
    src_bf16 = "MEASURED_PEAKS.json bf16_tflops" if peaks else "fallback 1590 (B200_PROFILING.md)"
    try:
        tops, at = 0.0, 0
        for n in (8192, 16384):  # the larger problem amortises launch and tail effects: take the better one
            a = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=device)
            b = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=device)
            for _ in range(3):
                torch._int_mm(a, b)
            best = 1e9
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                torch._int_mm(a, b)
                e1.record()
                e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            t = 2 * n ** 3 / (best * 1e-3) / 1e12
            if t > tops:
                tops, at = t, n
            del a, b
        return tops, bf16, f"torch._int_mm {at}^3 int8 (cuBLASLt), best of 10 over 8192^3 and 16384^3, measured in this run; bf16 from {src_bf16}"
    except Exception as ex:  # noqa: BLE001
        return 2 * bf16, bf16, f"2 x {src_bf16} (torch._int_mm unavailable: {type(ex).__name__})"


def main_ours(args):
    import torch
    import torch.distributed as dist

    from secedo_b200 import api
    from secedo_b200 import dist as sdist
    from secedo_b200.pileup import Pileup

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: secedo_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    w = WORKLOAD
    N = w["n_cells"]
    ctx = api.Context(local_rank)
    # one explicit stream for everything: the library's kernels, the NCCL reduce and the timing events
    stream = torch.cuda.Stream(device)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ident = np.arange(N, dtype=np.uint32)

    raw_dev = ctx.synth_pileup(N, w["coverage"], w["n_chr"], w["loci_per_chr"], n_clones=w["n_clones"],
                               frac_somatic=w["frac_somatic"], frac_germline=w["frac_germline"], theta=w["theta"],
                               spacing=w["spacing"], p_multi=w["p_multi"], p_mate=w["p_mate"],
                               p_mate_mismatch=w["p_mate_mismatch"], seed=1000 + rank)
    n_chr, P, E = raw_dev.dims()
    flt = api.Filter(w["theta"], 4, ctx)
    counts = api.Counts(ctx, N)
    lik = (w["L"], w["eps"], w["h"], w["theta"])
    last = {}

    def step(src, out_host):
        """one pass of the hot path; src is a DevicePileup (resident) or a host Pileup (e2e)"""
        filtered, cov = flt.filter_device(src, ident)
        counts.zero()
        st = counts.accumulate(filtered, w["L"], ident, w["eps"], w["h"], w["theta"], w["num_threads"], args.path)
        st["sig_loci"] = filtered.n_loci
        filtered.free()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record(stream)
        sdist.reduce_counts(counts, device, dst=0)
        r1.record(stream)
        if world > 1:
            r1.synchronize()
            st["ms_reduce"] = r0.elapsed_time(r1)
        if rank == 0:
            counts.finalize(*lik, w["normalization"], out=out_host, to_host=out_host is not None)
        last.update(st)
        return st

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stop, lines = threading.Event(), []
        th = threading.Thread(target=clocks_sampler, args=(stop, lines, local_rank), daemon=True)
        if rank == 0:
            th.start()
            time.sleep(0.25)  # let nvidia-smi come up BEFORE the barrier, so that all ranks start together
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        launches0 = ctx.launch_count()
        acc = {"ms_gemm": 0.0, "ms_stage": 0.0, "gemm_launches": 0, "ms_link": 0.0, "ms_first_order": 0.0,
               "ms_multi": 0.0, "ms_epilogue": 0.0, "ms_reduce": 0.0, "sig_loci": 0}
        e0.record(stream)
        for _ in range(steps):
            st = fn()
            for k in acc:
                acc[k] += st.get(k, 0) or 0
        e1.record(stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        stop.set()
        launches = ctx.launch_count() - launches0
        return float(ms.item()), acc, launches, summarize_clocks(lines)

    # ---- device-resident input ------------------------------------------------------------------------
    ms_dev, acc, launches, clocks = timed(lambda: step(raw_dev, None), args.steps, args.warmup)
    sig_local = acc["sig_loci"] // args.steps
    sig = torch.tensor([sig_local], device=device, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(sig)
    sig_total = int(sig.item())

    # ---- end to end: host (pinned) pileup in, host matrix out ------------------------------------------
    # The reference-facing call sequence with HOST buffers: every step uploads its whole batch (one
    # asynchronous upload per chromosome on the library's copy stream, so that the copy of the next
    # chromosome overlaps the kernels of the current one, and the first copy of the next step overlaps the
    # D2H of this step's matrix), filters and accumulates chromosome by chromosome into one counts object,
    # reduces, and reads the N x N matrix back into host memory. All H2D / D2H bytes of all timed steps
    # are inside the timed region.
    host = raw_dev.download()
    pinned = []
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        pinned.append(t)
        return t.numpy()
    pos_p, rid_p, gb_p = pin(host.position), pin(host.read_id), pin(host.gid_base)
    chunks = []
    for c in range(n_chr):
        l0, l1 = int(host.chr_ptr[c]), int(host.chr_ptr[c + 1])
        e0, e1 = int(host.row_ptr[l0]), int(host.row_ptr[l1])
        hp = Pileup.__new__(Pileup)  # views of the pinned arrays, no copies
        hp.chr_ptr = np.array([0, l1 - l0], np.uint64)
        hp.row_ptr = pin(host.row_ptr[l0:l1 + 1] - np.uint64(e0))
        hp.position, hp.read_id, hp.gid_base = pos_p[l0:l1], rid_p[e0:e1], gb_p[e0:e1]
        chunks.append(hp)
    out_t = torch.empty((N, N), dtype=torch.float64).pin_memory() if rank == 0 else None
    out_host = out_t.numpy() if rank == 0 else None
    # second result buffer: the matrix of step s travels to the host (own stream) while step s + 1 is computed
    out_t2 = torch.empty((N, N), dtype=torch.float64).pin_memory() if rank == 0 else None
    out_bufs = [out_host, out_t2.numpy()] if rank == 0 else [None, None]
    e2e_steps = max(1, min(args.steps, 3))
    E2E_WARMUP = 2
    e2e_total = E2E_WARMUP + e2e_steps  # warm-up + timed
    # uploads run up to LOOKAHEAD chromosomes ahead of the kernels, but never across the warm-up / timed
    # boundary nor past the last step: every timed step's H2D bytes are copied inside the timed region
    # (a whole step ahead: the DMA engine then works through accumulate and the D2H of the matrix, when the kernels
    # leave the H2D direction of the bus idle, and the filter's zero-copy pull has the bus to itself)
    LOOKAHEAD = 4
    runs = [[(s, c) for s in range(E2E_WARMUP) for c in range(n_chr)],
            [(s, c) for s in range(E2E_WARMUP, e2e_total) for c in range(n_chr)]]
    state = {"queue": [], "issued": 0, "run": []}

    def top_up(target):
        while len(state["queue"]) < target and state["issued"] < len(state["run"]):
            # read ids stay in pinned host memory: the filter pulls those of the loci it keeps (zero copy)
            state["queue"].append(ctx.upload_lazy_async(chunks[state["run"][state["issued"]][1]]))
            state["issued"] += 1

    def e2e_step():
        if state["issued"] == len(state["run"]) and not state["queue"]:  # next run (warm-up, then the timed steps)
            state["run"], state["issued"] = runs.pop(0), 0
        counts.zero()
        st = {}
        tr = state.setdefault("trace", {}) if os.environ.get("SECEDO_BENCH_E2E_TRACE") else None
        t_prev = time.perf_counter()

        def lap(name):
            nonlocal t_prev
            if tr is not None:
                now = time.perf_counter()
                tr[name] = tr.get(name, 0.0) + (now - t_prev) * 1e3
                t_prev = now
        for c in range(n_chr):
            top_up(1)
            cur = state["queue"].pop(0)
            lap("issue")
            filtered, _ = flt.filter_device(cur, ident)
            lap("filter")
            # DMA of the coming chromosomes is queued where the kernels leave the H2D direction of the bus idle
            # (accumulate, and the D2H of the matrix below), not next to the filter's zero-copy pull
            top_up(2)
            lap("issue")
            s1 = counts.accumulate(filtered, w["L"], ident, w["eps"], w["h"], w["theta"], w["num_threads"], args.path)
            st["sig_loci"] = st.get("sig_loci", 0) + filtered.n_loci
            state["pulled_entries"] = state.get("pulled_entries", 0) + filtered.n_entries
            for k in ("ms_gemm", "ms_stage", "ms_link", "ms_first_order", "ms_multi", "gemm_launches"):
                st[k] = st.get(k, 0) + (s1.get(k, 0) or 0)
            lap("accumulate")
            filtered.free()
            cur.free()
            lap("free")
        sdist.reduce_counts(counts, device, dst=0)
        top_up(LOOKAHEAD)
        lap("issue")
        if rank == 0:
            # the previous step's matrix has to be complete before its buffer's turn comes again; every matrix of
            # the timed steps is complete before the timed region ends (output_wait below, inside `timed`'s last step)
            state["n_out"] = state.get("n_out", 0) + 1
            counts.finalize_async(*lik, w["normalization"], out_bufs[state["n_out"] % 2])
            if state["n_out"] in (E2E_WARMUP, e2e_total):  # last step of a run: nothing left to overlap with
                ctx.output_wait()
        lap("finalize")
        if tr is not None:
            sys.stderr.write("e2e trace (ms, cumulative): " + json.dumps({k: round(v, 2) for k, v in tr.items()}) + "\n")
        return st

    ms_e2e, acc_e2e, _, _ = timed(e2e_step, e2e_steps, E2E_WARMUP)
    ctx.synchronize()
    if rank == 0:
        ctx.output_wait()
        out_host = out_bufs[state["n_out"] % 2]
    # bytes that crossed PCIe host -> device per step: the CSR without the read ids, plus the read ids of the
    # entries of the loci the filter kept (all cells take part: every entry of a kept locus is pulled once)
    pulled = state.get("pulled_entries", 0) // e2e_total
    h2d_full = host.row_ptr.nbytes + host.position.nbytes + host.read_id.nbytes + host.gid_base.nbytes + host.chr_ptr.nbytes
    h2d = host.row_ptr.nbytes + host.position.nbytes + host.gid_base.nbytes + host.chr_ptr.nbytes + 4 * pulled
    d2h = N * N * 8
    sig_e2e = torch.tensor([acc_e2e["sig_loci"] // e2e_steps], device=device, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(sig_e2e)
    assert int(sig_e2e.item()) == sig_total, "the chromosome-wise end-to-end path must see the same significant loci"

    if rank == 0:
        M = out_host
        assert np.array_equal(M, M.T) and not np.diag(M).any(), "result must be symmetric with a zero diagonal"

    # ---- roofline of the dominant kernel (syrk_kernel, tensor bound) ---------------------------------------
    ms_gemm_step = acc["ms_gemm"] / args.steps
    launches_gemm = max(1, acc["gemm_launches"] // args.steps)
    alg_ops_step = 5.0 * N * N * sig_local                      # 5 N^2 per significant locus
    n_pad = ((N + 255) // 256) * 256
    n_tiles = sum(1 for cb in range(n_pad // 256) for rb in range(n_pad // 128)
                  if cb * 256 + 255 > rb * 128 and rb * 128 < N and cb * 256 < N)
    kbs = (sig_local + 31) // 32
    executed_ops_step = 2.0 * n_tiles * 128 * 256 * kbs * 128  # MACs x 2 the tcgen05 kernel issues: tiles x k-blocks of 128 B
    line = None
    if rank == 0:
        int8_peak, bf16_peak, how = int8_peak_tops(torch, device)
        achieved = alg_ops_step / (ms_gemm_step * 1e-3) / 1e12 if ms_gemm_step > 0 else 0.0
        # DRAM bytes of one launch from the committed ncu --set full capture of this exact workload
        traffic, tensor_active = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_syrk_traffic.json")))
            tw = tr["workload"]
            if (tw["n_cells"], tw["n_chr"], tw["loci_per_chr"], tw["coverage"]) == (N, w["n_chr"], w["loci_per_chr"], w["coverage"]):
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
                tensor_active = tr.get("tensor_pipe_active_pct")
        except Exception:  # noqa: BLE001
            pass
        NOMINAL_INT8 = 4500.0  # dense int8 TOP/s of a B200 (B200_PROFILING.md: 2 x the 2.25 PFLOP/s bf16 figure)
        roofline = {
            "bound": "tensor", "kernel": "syrk2_kernel (tcgen05.mma.cta_group::2.kind::i8)", "achieved": achieved, "peak": int8_peak,
            "unit": "TFLOP/s", "frac": achieved / int8_peak if int8_peak else None, "traffic": traffic,
            "op": "int8 multiply-add = 2 ops; algorithmic ops = 5*N^2 per significant locus (SURVEY 8d); the kernel "
                  "issues 4*N_pad^2*(upper-triangle tiles) of them (Hadamard planes: 4 K-slices per 32 loci instead of 5)",
            "peak_source": how, "bf16_peak_measured": bf16_peak,
            "launches_per_step": launches_gemm, "avg_launch_ms": ms_gemm_step / launches_gemm,
            "executed_ops_frac_of_algorithmic": executed_ops_step / alg_ops_step if alg_ops_step else None,
            "achieved_executed": achieved * executed_ops_step / alg_ops_step if alg_ops_step else None,
            "frac_executed": (achieved * executed_ops_step / alg_ops_step / int8_peak) if alg_ops_step and int8_peak else None,
            "kernel_share_of_step": ms_gemm_step / (ms_dev / args.steps),
            "nominal_peak": NOMINAL_INT8,
            "frac_executed_of_nominal": (achieved * executed_ops_step / alg_ops_step / NOMINAL_INT8) if alg_ops_step else None,
            "tensor_pipe_active_pct_ncu": tensor_active,
            "note": "frac > 1: `achieved` counts SURVEY's algorithmic ops, the Hadamard form issues 0.86 of them; and the "
                    "measured cuBLASLt int8 peak (like the bf16 one in MEASURED_PEAKS.json, 74 % of nominal) is a long "
                    "power-capped GEMM loop, while this kernel runs for a few ms between memory-bound kernels. In issued "
                    "ops it reaches frac_executed_of_nominal of the 4.5 POP/s dense int8 rate, tensor pipe active as "
                    "measured by ncu in tensor_pipe_active_pct_ncu",
            "traffic_source": "profiles/r1_syrk_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)",
        }
        # ---- Laplacian + the 7 leading eigenpairs of this step's matrix, matrix resident in HBM --
        spectral = None
        try:
            runs = []
            for _ in range(3):
                t0 = time.perf_counter()
                ev, _, sst, _ = counts.finalize_spectral(*lik, w["normalization"], k=7, tol=1e-10)
                sst["wall_ms"] = (time.perf_counter() - t0) * 1e3
                runs.append(sst)
            best = min(runs[1:], key=lambda r: r["wall_ms"])
            hbm = 6650.0
            try:
                hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            except Exception:  # noqa: BLE001
                pass
            gbs = 8.0 * N * N * best["matvec_launches"] / (best["ms_matvec"] * 1e-3) / 1e9
            spectral = {"what": "epilogue + laplacian() + the 7 smallest eigenpairs (spectral_clustering.cpp:33-52,127-138), "
                                "similarity matrix never leaves HBM; residuals <= 1e-10", "k": 7,
                        "ms": best["wall_ms"], "ms_solver": best["ms_solver"], "ms_laplacian": best["ms_laplacian"],
                        "outer_iterations": best["outer_iterations"], "block": best["block"],
                        "block_products": best["matvec_launches"], "max_residual": best["max_residual"],
                        "eigenvalues": [float(x) for x in ev],
                        "roofline": {"bound": "hbm", "kernel": "symm_block_kernel (fp64, M read once per product)",
                                     "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                                     "avg_launch_ms": best["ms_matvec"] / best["matvec_launches"],
                                     "algorithmic_bytes_per_launch": 8.0 * N * N}}
        except Exception as ex:  # noqa: BLE001
            spectral = {"error": f"{type(ex).__name__}: {ex}"}
        # ---- EM refinement on this rank's filtered batch (device resident) -------------------
        em = None
        try:
            from oracle import pyoracle as po
            filtered, _ = flt.filter_device(raw_dev, ident)
            start = np.random.default_rng(1).uniform(0.3, 0.7, N)
            EM_IT = 4
            api.expectation_maximization(filtered, ident, 1, w["theta"], start, ctx=ctx, max_iterations=1)  # warm
            _, est = api.expectation_maximization(filtered, ident, 1, w["theta"], start, ctx=ctx, max_iterations=EM_IT,
                                                  return_stats=True)
            n_ent = filtered.n_entries
            filtered.free()
            gpu_eps = n_ent * est["iterations"] / (est["ms"] * 1e-3)
            em = {"what": "expectation_maximization (expectation_maximization.cpp:131-160) on the filtered batch in HBM, "
                          f"{EM_IT} iterations incl. H2D/D2H of the probabilities", "entries": int(n_ent),
                  "ms_per_iteration": est["ms"] / est["iterations"], "entries_per_s": gpu_eps,
                  "algorithmic_bytes_per_entry": 4, "achieved_GBps": 4.0 * gpu_eps / 1e9,
                  "bound": "L2 reductions (one RED.ADD.64 per entry) + two dependent gathers per entry"}
            sp_cfg_loci = 256
            from secedo_b200.synth import SynthConfig, make_pileup
            scfg = SynthConfig(n_cells=N, coverage=w["coverage"], n_loci=sp_cfg_loci, n_chr=1, n_clones=w["n_clones"],
                               frac_somatic=0.5, frac_germline=0.0, theta=w["theta"], spacing=w["spacing"], seed=5)
            sp = make_pileup(scfg)
            if po.have_ref():
                _, it_ref = po.expectation_maximization(sp, ident, w["theta"], start)
                _, secs = po.ref_expectation_maximization(sp, ident, w["theta"], start)
                em["cpu_reference"] = {"entries": int(sp.n_entries), "iterations": it_ref, "seconds": secs, "cores": 1,
                                       "entries_per_s": sp.n_entries * it_ref / secs,
                                       "note": "the reference's EM is serial (its omp pragma is commented out, "
                                               "expectation_maximization.cpp:140)"}
        except Exception as ex:  # noqa: BLE001
            em = {"error": f"{type(ex).__name__}: {ex}"}
        # ---- CPU baseline: the unmodified reference on a bounded sample ---------------------------------
        threads = reference_threads()
        sample_p = cpu_sample()
        secs, cpu_loci, kind = run_reference_step(sample_p, threads)
        cpu_value = cpu_loci / secs
        cpu = {"value": cpu_value, "unit": "loci/s", "cores": threads, "kind": kind,
               "sample": f"{CPU_SAMPLE_LOCI} pre-filter loci ({cpu_loci} significant, {sample_p.n_entries} entries) of the "
                         f"same workload, Filter::filter + computeSimilarityMatrix, {secs:.1f} s"}
        value = sig_total * args.steps / (ms_dev * 1e-3)
        e2e_value = sig_total * e2e_steps / (ms_e2e * 1e-3)
        line = {
            "metric": "similarity-matrix significant loci/s (8K cells, 0.5x)", "value": value, "unit": "loci/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8 tensor-core counts (int32 accumulate) + f64 log-likelihood epilogue", "data": "synthetic",
            "config": {
                "workload": w["name"] % w["loci_per_chr"], "n_cells": N, "coverage": w["coverage"],
                "prefilter_loci_per_gpu_step": P, "significant_loci_per_step_all_gpus": sig_total,
                "pileup_entries_per_gpu_step": E, "p_multi": w["p_multi"], "p_mate": w["p_mate"],
                "theta": w["theta"], "eps": w["eps"], "h": w["h"], "max_fragment_length": w["L"],
                "num_threads_for_cutoff": w["num_threads"], "normalization": w["normalization"],
                "path": last.get("path_used"), "parallelism": f"loci sharded by chromosome over {world} GPU(s), "
                "one NCCL reduce of the int32 count planes in use",
                "l2": "inputs larger than L2 (pileup batch %.1f GB, Hadamard panel %.1f GB per step)" % (
                    h2d_full / 1e9, 4.0 * N * sig_local / 1e9),
            },
            "build_time_s_per_step": ms_dev / args.steps * 1e-3,
            "prefilter_loci_per_s": P * world * args.steps / (ms_dev * 1e-3),
            "phase_ms_per_step_rank0": {k: acc[k] / args.steps for k in ("ms_link", "ms_first_order", "ms_stage", "ms_gemm",
                                                                         "ms_multi", "ms_epilogue", "ms_reduce")},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "loci/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
                    "host_pileup_bytes_per_step": int(h2d_full),
                    "how": "host pinned pileup -> sgpu_pileup_upload_lazy_async per chromosome (copy stream, overlapping the "
                           "kernels of the previous chromosome; the read ids stay in pinned host memory) -> filter (pulls the "
                           "read ids of the loci it keeps straight over PCIe: h2d_bytes_per_step counts what crossed the "
                           "bus, host_pileup_bytes_per_step the whole input) -> accumulate -> reduce -> finalize -> N x N "
                           "fp64 matrix in host memory, every step"},
            "gpu_launches": int(launches), "clocks": clocks, "spectral": spectral, "em": em,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--path", default="auto", choices=["auto", "scatter", "gemm"])
    a = ap.parse_args()
    sys.exit(main_reference(a) if a.impl == "reference" else main_ours(a))
