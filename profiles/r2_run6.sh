#!/bin/bash
# round 2, GPU call 6 (2 GPUs): full gpu suite incl. single-process multi-GPU (sgpu_multi_*), shim over 1 and 2 GPUs; cfg5 crossover line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_pytest6.log
timeout 900 python profiles/path_crossover.py > gpurun_out/r2_path_crossover.txt 2> gpurun_out/r2_path_crossover.err; echo "crossover rc=$?"
