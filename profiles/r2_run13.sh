#!/bin/bash
# round 2, GPU call 13 (2 GPUs): full gpu suite after the last kernel changes (sp_finish, significance test restructured), bench N = 1, whole genome N = 2 at its real size
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest13.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_pytest13.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench13_n1.json 2> gpurun_out/r2_bench13_n1.err; echo "bench n1 rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
SECEDO_BENCH_GENOME_LOCI=16300000 timeout 900 $TR --nproc-per-node 2 --master-port 29552 bench.py --workload cfg3-genome --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2_genomefull_n2.json 2> gpurun_out/r2_genomefull_n2.err; echo "genome full n2 rc=$?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 python profiles/hot_step.py 3 > gpurun_out/hot_plain.log 2>&1 && \
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches13.csv python profiles/hot_step.py 3 > gpurun_out/ncu13.log 2>&1
echo "ncu launches rc=$?"
tail -n 3 gpurun_out/r2_bench13_n1.err gpurun_out/r2_genomefull_n2.err
