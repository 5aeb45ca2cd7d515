#!/bin/bash
# round 2, GPU call 1: full gpu test suite, baseline bench, launch list with DRAM bytes, GEMM raster / L2-promotion probes
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_env.txt; nproc >> gpurun_out/r2_env.txt; free -g >> gpurun_out/r2_env.txt
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest1.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
python profiles/hot_step.py 3 > gpurun_out/hot_plain.log 2>&1 && \
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches1.csv python profiles/hot_step.py 3 > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
for cfg in "4 256" "16 256" "31 256" "8 128" "8 0"; do
  set -- $cfg
  SECEDO_B200_TILE_BAND=$1 SECEDO_B200_L2_PROMO=$2 ncu -k regex:syrk2 --metrics $M --clock-control none --csv \
     --log-file gpurun_out/r2_raster_band$1_promo$2.csv python profiles/hot_step.py 2 > gpurun_out/ncu_raster.log 2>&1
  echo "raster $cfg rc=$?"; grep -c syrk2 gpurun_out/r2_raster_band$1_promo$2.csv
done
