#!/bin/bash
# round 2, GPU call 33: last check of the library as committed (cache preference restored at shutdown): the scheduling-knob test,
# the interleaved-objects test, smoke, a short bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_similarity.py -m gpu -x -q -k "scheduling_knobs or two_counts or view or auto_path" > gpurun_out/r2_pytest33.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_pytest33.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2_smoke33.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r2_smoke33.log
timeout 600 python bench.py --steps 6 --warmup 3 --skip-extras > gpurun_out/r2_bench33_n1.json 2> gpurun_out/r2_bench33_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench33_n1.json").read().strip().splitlines()[-1])
print("bench: value=%.4g ms/step=%.2f e2e=%.1f parity=%s clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity_vs_reference"]["ok"], d["clocks"]))
PY
