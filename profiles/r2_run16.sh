#!/bin/bash
# round 2, GPU call 16: why did the kernels of the next sub-batch not run beside the tensor kernel? co-residency probe
mkdir -p gpurun_out
timeout 120 ./profiles/coresidency_probe > gpurun_out/r2_coresidency16.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r2_coresidency16.txt
P="timeout 300 python profiles/overlap_probe.py 6 2"
: > gpurun_out/r2_overlap_probe16.txt
run() { label=$1; shift; env "$@" $P "$label" >> gpurun_out/r2_overlap_probe16.txt 2>> gpurun_out/r2_overlap_probe16.err || echo "probe $label failed"; }
run async_s5_full_prefshared  SECEDO_B200_ASYNC_GEMM=1 SECEDO_B200_GEMM_STAGES=5 SECEDO_B200_WIN_SMEM_KB=220 SECEDO_B200_PREFER_SHARED=1
run async_s5_win60_prefshared SECEDO_B200_ASYNC_GEMM=1 SECEDO_B200_GEMM_STAGES=5 SECEDO_B200_WIN_SMEM_KB=60 SECEDO_B200_PREFER_SHARED=1
python - <<'PY'
import json
for l in open("gpurun_out/r2_overlap_probe16.txt"):
    d = json.loads(l)
    print("%-26s %7.2f ms/step  %6.2f /sub  gemm %.2f ms x %d  phases %s chk %x" % (d["label"], d["ms_per_step"], d["ms_per_sub_batch"], d["gemm_avg_ms"], d["phase_ms_per_step"]["gemm_launches"], {k: round(v, 2) for k, v in d["phase_ms_per_step"].items() if k != "gemm_launches"}, d["checksum"]))
PY
tail -n 5 gpurun_out/r2_overlap_probe16.err
