#!/bin/bash
# round 2, GPU call 8 (2 GPUs): full gpu suite; cfg5 bench line; whole-genome mode (segments, ranged accumulation) at N = 1, 2
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest8.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_pytest7.log
timeout 900 python bench.py --workload cfg5 --steps 5 --warmup 3 > gpurun_out/r2_bench8_cfg5.json 2> gpurun_out/r2_bench8_cfg5.err; echo "cfg5 rc=$?"
timeout 600 python bench.py --workload cfg3-genome --steps 3 --warmup 1 > gpurun_out/r2_genome8_n1.json 2> gpurun_out/r2_genome8_n1.err; echo "genome n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload cfg3-genome --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2_genome8_n2.json 2> gpurun_out/r2_genome8_n2.err; echo "genome n2 rc=$?"
tail -n 4 gpurun_out/r2_bench8_cfg5.err gpurun_out/r2_genome8_n1.err gpurun_out/r2_genome8_n2.err
