#!/bin/bash
# usage: quick_bench.sh [label]  — short bench run, prints the numbers that matter
python bench.py --steps 5 --warmup 3 > gpurun_out/qb_$1.json 2> gpurun_out/qb_$1.err || { echo "bench failed"; tail -5 gpurun_out/qb_$1.err; exit 1; }
python - "$1" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/qb_{sys.argv[1]}.json"))
ph = {k: round(v, 2) for k, v in d["phase_ms_per_step_rank0"].items()}
print(sys.argv[1], "loci/s=%.0f ms/step=%.2f" % (d["value"], d["ms_per_step"]), ph, "e2e ms=%.1f" % d["e2e"]["ms_per_step"],
      "frac=%.3f" % d["roofline"]["frac"], "launches=%d" % d["gpu_launches"], d["clocks"])
PY
