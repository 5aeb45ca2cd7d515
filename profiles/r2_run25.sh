#!/bin/bash
# round 2, GPU call 25: per-object tensor joins; bench steps pipelined over two sets of count planes (the last tensor kernel of a
# matrix beside the next matrix's first sub-batch)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_similarity.py tests/test_gpu_pieces.py -m gpu -x -q -k "several_gemm or auto_path or int8 or view or second_order or pieces_match or device_generated" > gpurun_out/r2_pytest25.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_pytest25.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench25_pipelined.json 2> gpurun_out/r2_bench25_pipelined.err; echo "bench pipelined rc=$?"
SECEDO_BENCH_PIPELINE_STEPS=0 timeout 600 python bench.py --steps 8 --warmup 3 --skip-extras > gpurun_out/r2_bench25_unpipelined.json 2> gpurun_out/r2_bench25_unpipelined.err; echo "bench unpipelined rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench25_pipelined.json", "gpurun_out/r2_bench25_unpipelined.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, "value=%.4g ms/step=%.2f e2e=%.1f parity=%s gemm live %.2f alone %s share %.2f" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity_vs_reference"]["ok"], r["avg_launch_ms"], r["avg_launch_ms_alone"], r["kernel_share_of_step"]))
        print("   phases", d["phase_ms_per_step_rank0"])
        if d.get("e2e_shim"): print("   e2e_shim", {k: d["e2e_shim"].get(k) for k in ("significant_loci", "first_call_s", "repeat_call_s", "loci_per_s_first", "loci_per_s_repeat", "error")})
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -n 3 gpurun_out/r2_bench25_pipelined.err
