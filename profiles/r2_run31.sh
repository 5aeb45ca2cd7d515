#!/bin/bash
# round 2, GPU call 31 (4 GPUs): default bench at N = 4 in the final state (scaling table)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 4 --master-port 29541 bench.py --gpus 4 --steps 6 --warmup 3 > gpurun_out/r2_bench31_n4.json 2> gpurun_out/r2_bench31_n4.err; echo "bench n4 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench31_n4.json").read().strip().splitlines()[-1])
r = d.get("roofline") or {}
print("value=%.4g ms/step=%.2f e2e=%s parity=%s chk=%s gemm live %.2f alone %s" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("ms_per_step"), (d.get("parity_vs_reference") or {}).get("ok"), d.get("reduce_checksum_ok"), r.get("avg_launch_ms", 0), r.get("avg_launch_ms_alone")))
PY
tail -n 3 gpurun_out/r2_bench31_n4.err
