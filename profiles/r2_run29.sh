#!/bin/bash
# round 2, GPU call 29: final state — full gpu suite, smoke, bench N = 1, ncu captures of the non-tensor kernels of the hot step and of
# the pair-scatter kernel (ultra-sparse workload: 8 000 cells at 0.005x, where SGPU_PATH_AUTO takes the scatter path)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest29.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_pytest29.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2_smoke29.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/r2_smoke29.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench29_n1.json 2> gpurun_out/r2_bench29_n1.err; echo "bench n1 rc=$?"
timeout 600 ncu -k "regex:filter_count|filter_compact|partition_kernel|stage_tile|multilocus_kernel|mate_rule" --set full --clock-control none --import-source on -s 6 -c 6 -o gpurun_out/r2_hot_kernels_full python profiles/hot_step.py 2 > gpurun_out/ncu29a.log 2>&1; echo "ncu hot kernels rc=$?"
SECEDO_BENCH_COVERAGE=0.005 timeout 600 ncu -k regex:scatter_pairs --set full --clock-control none --import-source on -s 1 -c 1 -o gpurun_out/r2_scatter_full python profiles/hot_step.py 2 > gpurun_out/ncu29b.log 2>&1; echo "ncu scatter rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench29_n1.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("bench: value=%.4g ms/step=%.2f e2e=%.1f parity=%s gemm live %.2f alone %.2f share %.2f launches %d clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity_vs_reference"]["ok"], r["avg_launch_ms"], r["avg_launch_ms_alone"], r["kernel_share_of_step"], d["gpu_launches"], d["clocks"]))
PY
tail -n 3 gpurun_out/r2_bench29_n1.err gpurun_out/ncu29a.log gpurun_out/ncu29b.log
