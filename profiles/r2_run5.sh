#!/bin/bash
# round 2, GPU call 5 (2 GPUs): full gpu suite (pieces, wide ids, span splitting, shim cache, peer-memory epilogue with spill), path crossover
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_pytest5.log
timeout 900 python profiles/path_crossover.py > gpurun_out/r2_path_crossover.txt 2> gpurun_out/r2_path_crossover.err; echo "crossover rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench5_n1.json 2> gpurun_out/r2_bench5_n1.err; echo "bench n1 rc=$?"
