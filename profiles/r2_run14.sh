#!/bin/bash
# round 2, GPU call 14 (2 GPUs): full gpu suite after the ingest error-path fix
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest14.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_pytest14.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke14.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/r2_smoke14.log
