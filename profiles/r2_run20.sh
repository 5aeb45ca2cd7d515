#!/bin/bash
# round 2, GPU call 20: partition without the group table at the root (identity map), staging tile size beside the tensor kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_similarity.py tests/test_gpu_wide.py -m gpu -x -q -k "not full_cell_count and not huge_loci and not golden_matrices and not full_size" > gpurun_out/r2_pytest20.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2_pytest20.log
P="timeout 300 python profiles/overlap_probe.py 6 2"
: > gpurun_out/r2_overlap_probe20.txt
run() { label=$1; shift; env "$@" $P "$label" >> gpurun_out/r2_overlap_probe20.txt 2>> gpurun_out/r2_overlap_probe20.err || echo "probe $label failed"; }
run default
run cells320         SECEDO_B200_STAGE_CELLS=320
run cells256         SECEDO_B200_STAGE_CELLS=256
run cells208         SECEDO_B200_STAGE_CELLS=208
run s4_cells384      SECEDO_B200_GEMM_STAGES=4 SECEDO_B200_STAGE_CELLS=384
run sync_s6          SECEDO_B200_ASYNC_GEMM=0 SECEDO_B200_GEMM_STAGES=6
run sync_s6_cells256 SECEDO_B200_ASYNC_GEMM=0 SECEDO_B200_GEMM_STAGES=6 SECEDO_B200_STAGE_CELLS=256
python - <<'PY'
import json
for l in open("gpurun_out/r2_overlap_probe20.txt"):
    d = json.loads(l)
    print("%-18s %7.2f ms/step  %6.2f /sub  gemm %.2f ms x %d  phases %s chk %x" % (d["label"], d["ms_per_step"], d["ms_per_sub_batch"], d["gemm_avg_ms"], d["phase_ms_per_step"]["gemm_launches"], {k: round(v, 2) for k, v in d["phase_ms_per_step"].items() if k != "gemm_launches"}, d["checksum"]))
PY
tail -n 5 gpurun_out/r2_overlap_probe20.err
