#!/bin/bash
# round 2, GPU call 28 (8 GPUs): default bench at N = 8 with the overlapped tensor kernel and pipelined steps; 8-rank parity check
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/r2_bench28_n8.json 2> gpurun_out/r2_bench28_n8.err; echo "bench n8 rc=$?"
timeout 400 $TR --nproc-per-node 8 --master-port 29543 tests/multi_gpu_check.py > gpurun_out/r2_multi_check28_n8.log 2>&1; echo "multi check n8 rc=$?"; grep -c OK gpurun_out/r2_multi_check28_n8.log
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench28_n8.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, "value=%.4g ms/step=%.2f e2e=%s parity=%s chk=%s gemm live %.2f alone %s share %.2f" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("ms_per_step"), (d.get("parity_vs_reference") or {}).get("ok"), d.get("reduce_checksum_ok"), r.get("avg_launch_ms", 0), r.get("avg_launch_ms_alone"), r.get("kernel_share_of_step", 0)))
        print("   phases", d.get("phase_ms_per_step_rank0"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -n 4 gpurun_out/r2_bench28_n8.err; grep -n "Error" gpurun_out/r2_multi_check28_n8.log | head -3
