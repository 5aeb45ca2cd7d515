#!/bin/bash
# round 2, GPU call 34: the other workloads of bench.py in the final state (cfg5: 20 000 cells, wide ids; whole genome, default size)
mkdir -p gpurun_out
timeout 500 python bench.py --workload cfg5 --steps 4 --warmup 2 > gpurun_out/r2_bench34_cfg5.json 2> gpurun_out/r2_bench34_cfg5.err; echo "cfg5 rc=$?"
timeout 300 python bench.py --workload cfg3-genome --steps 2 --warmup 1 > gpurun_out/r2_bench34_genome.json 2> gpurun_out/r2_bench34_genome.err; echo "genome rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench34_cfg5.json", "gpurun_out/r2_bench34_genome.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, "value=%.4g ms/step=%.2f parity=%s gemm live %s alone %s" % (d["value"], d["ms_per_step"], (d.get("parity_vs_reference") or {}).get("ok"), r.get("avg_launch_ms"), r.get("avg_launch_ms_alone")))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -n 4 gpurun_out/r2_bench34_cfg5.err gpurun_out/r2_bench34_genome.err
