#!/bin/bash
# round 2, GPU call 9 (4 GPUs): default bench at N = 4; whole genome at its real size (16.3 M pre-filter loci) at N = 4 and N = 1
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/r2_env4.txt; nproc >> gpurun_out/r2_env4.txt; nvidia-smi topo -m >> gpurun_out/r2_env4.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 4 --master-port 29521 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_bench9_n4.json 2> gpurun_out/r2_bench9_n4.err; echo "bench n4 rc=$?"
SECEDO_BENCH_GENOME_LOCI=16300000 timeout 900 $TR --nproc-per-node 4 --master-port 29522 bench.py --workload cfg3-genome --gpus 4 --steps 3 --warmup 1 > gpurun_out/r2_genomefull_n4.json 2> gpurun_out/r2_genomefull_n4.err; echo "genome full n4 rc=$?"
SECEDO_BENCH_GENOME_LOCI=16300000 timeout 900 python bench.py --workload cfg3-genome --steps 1 --warmup 1 > gpurun_out/r2_genomefull_n1.json 2> gpurun_out/r2_genomefull_n1.err; echo "genome full n1 rc=$?"
tail -n 3 gpurun_out/r2_bench9_n4.err gpurun_out/r2_genomefull_n4.err gpurun_out/r2_genomefull_n1.err
