#!/bin/bash
# round 2, GPU call 30: staging kernel compiled for 4 / 3 / 2 CTAs per SM (32 / 40 / 56 registers; ncu's source view: the 32-register
# version recomputes addresses and spills in its output loop)
mkdir -p gpurun_out
P="timeout 300 python profiles/overlap_probe.py 6 2"
: > gpurun_out/r2_overlap_probe30.txt
run() { label=$1; shift; env "$@" $P "$label" >> gpurun_out/r2_overlap_probe30.txt 2>> gpurun_out/r2_overlap_probe30.err || echo "probe $label failed"; }
run minb4
run minb3            SECEDO_B200_STAGE_MINB=3
run minb2            SECEDO_B200_STAGE_MINB=2
run minb4_sync       SECEDO_B200_ASYNC_GEMM=0 SECEDO_B200_GEMM_STAGES=6
run minb3_sync       SECEDO_B200_ASYNC_GEMM=0 SECEDO_B200_GEMM_STAGES=6 SECEDO_B200_STAGE_MINB=3
run minb2_sync       SECEDO_B200_ASYNC_GEMM=0 SECEDO_B200_GEMM_STAGES=6 SECEDO_B200_STAGE_MINB=2
python - <<'PY'
import json
for l in open("gpurun_out/r2_overlap_probe30.txt"):
    d = json.loads(l)
    print("%-12s %7.2f ms/step  %6.2f /sub  gemm %.2f ms x %d  phases %s chk %x" % (d["label"], d["ms_per_step"], d["ms_per_sub_batch"], d["gemm_avg_ms"], d["phase_ms_per_step"]["gemm_launches"], {k: round(v, 2) for k, v in d["phase_ms_per_step"].items() if k != "gemm_launches"}, d["checksum"]))
PY
SECEDO_B200_STAGE_MINB=2 timeout 300 python -m pytest tests/test_gpu_similarity.py -m gpu -x -q -k "cfg1 or cfg2 or device_generated or several_gemm" > gpurun_out/r2_pytest30.log 2>&1; echo "pytest minb2 rc=$?"; tail -n 2 gpurun_out/r2_pytest30.log
tail -n 5 gpurun_out/r2_overlap_probe30.err
