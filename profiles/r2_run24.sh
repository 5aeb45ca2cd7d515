#!/bin/bash
# round 2, GPU call 24: bench with Filter::filter of the next sub-batch on a second context/host thread (SECEDO_BENCH_OVERLAP=1), now
# that kernels of different streams really share the SMs
mkdir -p gpurun_out
SECEDO_BENCH_OVERLAP=1 timeout 600 python bench.py --steps 8 --warmup 3 --skip-extras > gpurun_out/r2_bench24_filter_ahead.json 2> gpurun_out/r2_bench24_filter_ahead.err; echo "bench filter-ahead rc=$?"
SECEDO_BENCH_OVERLAP=1 SECEDO_B200_PREFER_SHARED=1 timeout 600 python bench.py --steps 8 --warmup 3 --skip-extras > gpurun_out/r2_bench24_filter_ahead_ps1.json 2> gpurun_out/r2_bench24_filter_ahead_ps1.err; echo "bench filter-ahead ps1 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench24_filter_ahead.json", "gpurun_out/r2_bench24_filter_ahead_ps1.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, "value=%.4g ms/step=%.2f e2e=%.1f gemm live %.2f alone %s share %.2f" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], r["avg_launch_ms"], r["avg_launch_ms_alone"], r["kernel_share_of_step"]))
        print("   phases", d["phase_ms_per_step_rank0"])
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -n 3 gpurun_out/r2_bench24_filter_ahead.err
