#!/bin/bash
# round 2, GPU call 17: co-residency needs cudaFuncCachePreferShared (call 16); which geometry lets the read linking run beside the tensor kernel?
mkdir -p gpurun_out
P="timeout 300 python profiles/overlap_probe.py 6 2"
: > gpurun_out/r2_overlap_probe17.txt
run() { label=$1; shift; env "$@" $P "$label" >> gpurun_out/r2_overlap_probe17.txt 2>> gpurun_out/r2_overlap_probe17.err || echo "probe $label failed"; }
A="SECEDO_B200_ASYNC_GEMM=1 SECEDO_B200_PREFER_SHARED=1"
run async_s5_full        $A SECEDO_B200_GEMM_STAGES=5 SECEDO_B200_WIN_SMEM_KB=220
run async_s6_full        $A SECEDO_B200_GEMM_STAGES=6 SECEDO_B200_WIN_SMEM_KB=220
run async_s4_full        $A SECEDO_B200_GEMM_STAGES=4 SECEDO_B200_WIN_SMEM_KB=220
run async_s5_win60_ring1 $A SECEDO_B200_GEMM_STAGES=5 SECEDO_B200_WIN_SMEM_KB=60 SECEDO_B200_WIN_RING=1
run async_s4_win60_ring1 $A SECEDO_B200_GEMM_STAGES=4 SECEDO_B200_WIN_SMEM_KB=60 SECEDO_B200_WIN_RING=1
run sync_s5_prefshared   SECEDO_B200_ASYNC_GEMM=0 SECEDO_B200_PREFER_SHARED=1 SECEDO_B200_GEMM_STAGES=5 SECEDO_B200_WIN_SMEM_KB=220
run sync_s5_ring1_60     SECEDO_B200_ASYNC_GEMM=0 SECEDO_B200_GEMM_STAGES=5 SECEDO_B200_WIN_SMEM_KB=60 SECEDO_B200_WIN_RING=1
python - <<'PY'
import json
for l in open("gpurun_out/r2_overlap_probe17.txt"):
    d = json.loads(l)
    print("%-22s %7.2f ms/step  %6.2f /sub  gemm %.2f ms x %d  phases %s chk %x" % (d["label"], d["ms_per_step"], d["ms_per_sub_batch"], d["gemm_avg_ms"], d["phase_ms_per_step"]["gemm_launches"], {k: round(v, 2) for k, v in d["phase_ms_per_step"].items() if k != "gemm_launches"}, d["checksum"]))
PY
tail -n 5 gpurun_out/r2_overlap_probe17.err
