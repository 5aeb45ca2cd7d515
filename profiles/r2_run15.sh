#!/bin/bash
# round 2, GPU call 15: tensor kernel on its own stream (overlaps the next sub-batch's filter / linking / staging):
# correctness subset, then A/B of the knobs on the device-resident step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_similarity.py tests/test_gpu_pieces.py -m gpu -x -q --durations=12 > gpurun_out/r2_pytest15.log 2>&1; echo "pytest rc=$?"; tail -n 18 gpurun_out/r2_pytest15.log
P="timeout 300 python profiles/overlap_probe.py 6 2"
: > gpurun_out/r2_overlap_probe15.txt
run() { label=$1; shift; env "$@" $P "$label" >> gpurun_out/r2_overlap_probe15.txt 2>> gpurun_out/r2_overlap_probe15.err || echo "probe $label failed"; }
run sync_s6_full   SECEDO_B200_ASYNC_GEMM=0 SECEDO_B200_GEMM_STAGES=6
run sync_s5_full   SECEDO_B200_ASYNC_GEMM=0 SECEDO_B200_GEMM_STAGES=5
run sync_s4_full   SECEDO_B200_ASYNC_GEMM=0 SECEDO_B200_GEMM_STAGES=4
run sync_s6_win60  SECEDO_B200_ASYNC_GEMM=0 SECEDO_B200_GEMM_STAGES=6 SECEDO_B200_WIN_SMEM_KB=60
run async_s6_win60 SECEDO_B200_ASYNC_GEMM=1 SECEDO_B200_GEMM_STAGES=6 SECEDO_B200_WIN_SMEM_KB=60
run async_s5_win60 SECEDO_B200_ASYNC_GEMM=1 SECEDO_B200_GEMM_STAGES=5 SECEDO_B200_WIN_SMEM_KB=60
run async_s4_win60 SECEDO_B200_ASYNC_GEMM=1 SECEDO_B200_GEMM_STAGES=4 SECEDO_B200_WIN_SMEM_KB=60
run async_s4_win92 SECEDO_B200_ASYNC_GEMM=1 SECEDO_B200_GEMM_STAGES=4 SECEDO_B200_WIN_SMEM_KB=92
run async_s5_full  SECEDO_B200_ASYNC_GEMM=1 SECEDO_B200_GEMM_STAGES=5 SECEDO_B200_WIN_SMEM_KB=220
python - <<'PY'
import json
for l in open("gpurun_out/r2_overlap_probe15.txt"):
    d = json.loads(l)
    print("%-16s %7.2f ms/step  %6.2f /sub  gemm %.2f ms x %d  phases %s chk %x" % (d["label"], d["ms_per_step"], d["ms_per_sub_batch"], d["gemm_avg_ms"], d["phase_ms_per_step"]["gemm_launches"], {k: round(v, 2) for k, v in d["phase_ms_per_step"].items() if k != "gemm_launches"}, d["checksum"]))
PY
tail -n 5 gpurun_out/r2_overlap_probe15.err
