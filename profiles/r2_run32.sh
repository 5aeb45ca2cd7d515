#!/bin/bash
# round 2, GPU call 32: ncu launch list of bench.py itself in the final state (gpu__time_duration per launch; cold-cache and
# serialised by ncu: shares, not absolutes), after the same command has run without ncu
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/r2_bench32_plain.json 2> gpurun_out/r2_bench32_plain.err; echo "plain rc=$?"
SECEDO_BENCH_CPU_LOCI=8 SECEDO_BENCH_E2E_STEPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_final.csv python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/ncu32.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r2_launches_bench_final.csv; tail -n 2 gpurun_out/ncu32.log | cut -c1-300
