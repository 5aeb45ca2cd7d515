#!/bin/bash
# round 2, GPU call 2: link_window v2 (contiguous owner ranges + cp.async ring), wave-synchronised GEMM, p_multi = 0.16 variant
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest2.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"
SECEDO_B200_WAVE_SYNC=0 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench2_nosync.json 2> gpurun_out/r2_bench2_nosync.err; echo "bench nosync rc=$?"
SECEDO_BENCH_P_MULTI=0.16 timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench2_pm16.json 2> gpurun_out/r2_bench2_pm16.err; echo "bench pm16 rc=$?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 python profiles/hot_step.py 3 > gpurun_out/hot_plain.log 2>&1 && \
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches2.csv python profiles/hot_step.py 3 > gpurun_out/ncu2.log 2>&1
echo "ncu launches rc=$?"
SECEDO_B200_WAVE_SYNC=0 timeout 600 ncu -k regex:syrk2 --metrics $M --clock-control none --csv --log-file gpurun_out/r2_syrk_nosync.csv python profiles/hot_step.py 2 > gpurun_out/ncu2b.log 2>&1
SECEDO_BENCH_P_MULTI=0.16 timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches2_pm16.csv python profiles/hot_step.py 2 > gpurun_out/ncu2c.log 2>&1
echo "ncu pm16 rc=$?"
