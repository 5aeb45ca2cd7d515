#!/bin/bash
# round 2, GPU call 19: the deferred tensor kernel is issued behind the NEXT batch's link_window kernel (late launch), device-wide
# cache preference toggled around the co-run window
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pieces.py tests/test_gpu_similarity.py -m gpu -x -q -k "not full_cell_count and not huge_loci and not golden_matrices" > gpurun_out/r2_pytest19.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r2_pytest19.log
P="timeout 300 python profiles/overlap_probe.py 6 2"
: > gpurun_out/r2_overlap_probe19.txt
run() { label=$1; shift; env "$@" $P "$label" >> gpurun_out/r2_overlap_probe19.txt 2>> gpurun_out/r2_overlap_probe19.err || echo "probe $label failed"; }
run default
run late_ps1_s5      SECEDO_B200_PREFER_SHARED=1
run late_ps2_s6      SECEDO_B200_GEMM_STAGES=6
run late_ps2_s4      SECEDO_B200_GEMM_STAGES=4
run early_ps1_s5     SECEDO_B200_GEMM_LATE=0 SECEDO_B200_PREFER_SHARED=1
run early_ps2_s5     SECEDO_B200_GEMM_LATE=0
run late_ps2_s5_w60  SECEDO_B200_WIN_SMEM_KB=60
run sync_s6          SECEDO_B200_ASYNC_GEMM=0 SECEDO_B200_GEMM_STAGES=6
python - <<'PY'
import json
for l in open("gpurun_out/r2_overlap_probe19.txt"):
    d = json.loads(l)
    print("%-18s %7.2f ms/step  %6.2f /sub  gemm %.2f ms x %d  phases %s chk %x" % (d["label"], d["ms_per_step"], d["ms_per_sub_batch"], d["gemm_avg_ms"], d["phase_ms_per_step"]["gemm_launches"], {k: round(v, 2) for k, v in d["phase_ms_per_step"].items() if k != "gemm_launches"}, d["checksum"]))
PY
tail -n 5 gpurun_out/r2_overlap_probe19.err
