#!/bin/bash
# round 2, GPU call 18: filtered pileups that read their ids through a view of the source (no copy of the read ids), phase
# timers without host waits, cudaFuncCachePreferShared per kernel instead of device-wide
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filter.py tests/test_gpu_em.py tests/test_shim.py tests/test_gpu_pipeline.py tests/test_gpu_similarity.py -m gpu -x -q -k "not full_size and not full_cell_count and not huge_loci and not golden_matrices" > gpurun_out/r2_pytest18.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r2_pytest18.log
P="timeout 300 python profiles/overlap_probe.py 6 2"
: > gpurun_out/r2_overlap_probe18.txt
run() { label=$1; shift; env "$@" $P "$label" >> gpurun_out/r2_overlap_probe18.txt 2>> gpurun_out/r2_overlap_probe18.err || echo "probe $label failed"; }
run sync_s6_copy         SECEDO_B200_ASYNC_GEMM=0 SECEDO_B200_GEMM_STAGES=6 SECEDO_B200_FILTER_VIEW=0
run sync_s6_view         SECEDO_B200_ASYNC_GEMM=0 SECEDO_B200_GEMM_STAGES=6
run async_s5_ps1_view    SECEDO_B200_ASYNC_GEMM=1 SECEDO_B200_GEMM_STAGES=5 SECEDO_B200_PREFER_SHARED=1
run async_s5_ps2_view    SECEDO_B200_ASYNC_GEMM=1 SECEDO_B200_GEMM_STAGES=5 SECEDO_B200_PREFER_SHARED=2
run async_s5_ps2_copy    SECEDO_B200_ASYNC_GEMM=1 SECEDO_B200_GEMM_STAGES=5 SECEDO_B200_PREFER_SHARED=2 SECEDO_B200_FILTER_VIEW=0
run async_s6_ps2_view    SECEDO_B200_ASYNC_GEMM=1 SECEDO_B200_GEMM_STAGES=6 SECEDO_B200_PREFER_SHARED=2
run default
python - <<'PY'
import json
for l in open("gpurun_out/r2_overlap_probe18.txt"):
    d = json.loads(l)
    print("%-20s %7.2f ms/step  %6.2f /sub  gemm %.2f ms x %d  phases %s chk %x" % (d["label"], d["ms_per_step"], d["ms_per_sub_batch"], d["gemm_avg_ms"], d["phase_ms_per_step"]["gemm_launches"], {k: round(v, 2) for k, v in d["phase_ms_per_step"].items() if k != "gemm_launches"}, d["checksum"]))
PY
tail -n 5 gpurun_out/r2_overlap_probe18.err
