#!/bin/bash
# round 2, GPU call 27 (2 GPUs): pipelined bench steps (two sets of count planes, two peer-memory epilogues) at N = 2, N-GPU == 1-GPU check
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29543 tests/multi_gpu_check.py > gpurun_out/r2_multi_check27.log 2>&1; echo "multi check n2 rc=$?"; grep -c OK gpurun_out/r2_multi_check27.log
timeout 600 $TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2_bench27_n2.json 2> gpurun_out/r2_bench27_n2.err; echo "bench n2 rc=$?"
timeout 600 $TR --nproc-per-node 2 --master-port 29542 bench.py --workload cfg3-genome --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2_genome27_n2.json 2> gpurun_out/r2_genome27_n2.err; echo "genome n2 rc=$?"
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_shim.py -m gpu -x -q > gpurun_out/r2_pytest27.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/r2_pytest27.log
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench27_n2.json", "gpurun_out/r2_genome27_n2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, "value=%.4g ms/step=%.2f e2e=%s parity=%s chk=%s gemm live %.2f alone %s share %.2f" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("ms_per_step"), (d.get("parity_vs_reference") or {}).get("ok"), d.get("reduce_checksum_ok"), r.get("avg_launch_ms", 0), r.get("avg_launch_ms_alone"), r.get("kernel_share_of_step", 0)))
        print("   phases", d.get("phase_ms_per_step_rank0"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -n 4 gpurun_out/r2_bench27_n2.err gpurun_out/r2_genome27_n2.err
