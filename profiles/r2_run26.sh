#!/bin/bash
# round 2, GPU call 26: staging tile of 4-bit counters (byte tile as the fallback)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_similarity.py tests/test_gpu_wide.py tests/test_gpu_pieces.py -m gpu -x -q -k "not golden_matrices and not cfg4" > gpurun_out/r2_pytest26.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2_pytest26.log
P="timeout 300 python profiles/overlap_probe.py 6 2"
: > gpurun_out/r2_overlap_probe26.txt
run() { label=$1; shift; env "$@" $P "$label" >> gpurun_out/r2_overlap_probe26.txt 2>> gpurun_out/r2_overlap_probe26.err || echo "probe $label failed"; }
run nibble
run bytes            SECEDO_B200_STAGE_NIBBLE=0
run nibble_sync_s6   SECEDO_B200_ASYNC_GEMM=0 SECEDO_B200_GEMM_STAGES=6
run bytes_sync_s6    SECEDO_B200_ASYNC_GEMM=0 SECEDO_B200_GEMM_STAGES=6 SECEDO_B200_STAGE_NIBBLE=0
run nibble_s6        SECEDO_B200_GEMM_STAGES=6
python - <<'PY'
import json
for l in open("gpurun_out/r2_overlap_probe26.txt"):
    d = json.loads(l)
    print("%-16s %7.2f ms/step  %6.2f /sub  gemm %.2f ms x %d  phases %s chk %x" % (d["label"], d["ms_per_step"], d["ms_per_sub_batch"], d["gemm_avg_ms"], d["phase_ms_per_step"]["gemm_launches"], {k: round(v, 2) for k, v in d["phase_ms_per_step"].items() if k != "gemm_launches"}, d["checksum"]))
PY
tail -n 5 gpurun_out/r2_overlap_probe26.err
