#!/bin/bash
# round 2, GPU call 3 (2 GPUs): partition v2, peer-memory epilogue (2-rank parity test), restructured bench at N=1 and N=2,
# whole-genome strong-scaling mode at N=1 and N=2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest3.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench3_n1.json 2> gpurun_out/r2_bench3_n1.err; echo "bench n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench3_n2.json 2> gpurun_out/r2_bench3_n2.err; echo "bench n2 rc=$?"
timeout 600 python bench.py --workload cfg3-genome --steps 3 --warmup 1 > gpurun_out/r2_genome_n1.json 2> gpurun_out/r2_genome_n1.err; echo "genome n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload cfg3-genome --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2_genome_n2.json 2> gpurun_out/r2_genome_n2.err; echo "genome n2 rc=$?"
tail -3 gpurun_out/r2_bench3_n1.err gpurun_out/r2_bench3_n2.err gpurun_out/r2_genome_n1.err gpurun_out/r2_genome_n2.err
