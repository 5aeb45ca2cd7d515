#!/bin/bash
# round 2, GPU call 12 (2 GPUs): A/B of the filter-ahead overlap at N = 1; 2-rank parity check with a lopsided assignment; launch list of the final kernels
mkdir -p gpurun_out
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench12_overlap.json 2> gpurun_out/r2_bench12_overlap.err; echo "bench overlap rc=$?"
SECEDO_BENCH_OVERLAP=0 timeout 600 python bench.py --steps 8 --warmup 3 --skip-extras > gpurun_out/r2_bench12_nooverlap.json 2> gpurun_out/r2_bench12_nooverlap.err; echo "bench no overlap rc=$?"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_pytest12.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_pytest12.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 python profiles/hot_step.py 3 > gpurun_out/hot_plain.log 2>&1 && \
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches12.csv python profiles/hot_step.py 3 > gpurun_out/ncu12.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu -k regex:syrk2 --set full --clock-control none --import-source on -s 1 -c 1 -o gpurun_out/r2_syrk2_full python profiles/hot_step.py 2 > gpurun_out/ncu12b.log 2>&1; echo "ncu syrk2 full rc=$?"
tail -n 3 gpurun_out/r2_bench12_overlap.err gpurun_out/r2_bench12_nooverlap.err
