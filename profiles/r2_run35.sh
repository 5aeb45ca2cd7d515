#!/bin/bash
# round 2, GPU call 35: the full gpu suite on the library as finally committed (clean rebuild)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest35.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_pytest35.log
