#!/bin/bash
# round 2, GPU call 11 (8 GPUs): default bench at N = 8, whole genome at its real size at N = 8, 8-rank parity check, single-process multi-GPU over 8 GPUs
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/r2_env8.txt; nproc >> gpurun_out/r2_env8.txt; nvidia-smi topo -m >> gpurun_out/r2_env8.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench11_n8.json 2> gpurun_out/r2_bench11_n8.err; echo "bench n8 rc=$?"
SECEDO_BENCH_GENOME_LOCI=16300000 timeout 600 $TR --nproc-per-node 8 --master-port 29542 bench.py --workload cfg3-genome --gpus 8 --steps 3 --warmup 1 > gpurun_out/r2_genomefull_n8.json 2> gpurun_out/r2_genomefull_n8.err; echo "genome full n8 rc=$?"
timeout 600 $TR --nproc-per-node 8 --master-port 29543 tests/multi_gpu_check.py > gpurun_out/r2_multi_check_n8.log 2>&1; echo "multi check n8 rc=$?"; grep -c OK gpurun_out/r2_multi_check_n8.log
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_shim.py -m gpu -x -q -k "single_process or all" > gpurun_out/r2_pytest11.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/r2_pytest11.log
tail -n 3 gpurun_out/r2_bench11_n8.err gpurun_out/r2_genomefull_n8.err
