#!/bin/bash
# round 2, GPU call 10 (2 GPUs): validation before the 8-GPU run: bench at N = 2 with the multi-GPU parity leg, new tests
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench10_n2.json 2> gpurun_out/r2_bench10_n2.err; echo "bench n2 rc=$?"
timeout 1200 python -m pytest tests/test_gpu_similarity.py tests/test_gpu_pieces.py tests/test_shim.py -m gpu -x -q > gpurun_out/r2_pytest10.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_pytest10.log
tail -n 5 gpurun_out/r2_bench10_n2.err
