#!/bin/bash
# round 2, GPU call 4 (2 GPUs): chromosomes in pieces (ranges + halo), spill planes in the peer-memory epilogue, link_window table-size probe
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_pytest4.log
M=gpu__time_duration.sum
for f in 1 2; do
SECEDO_B200_WIN_SLOT_FACTOR=$f timeout 600 ncu -k regex:link_window --metrics $M --clock-control none --csv --log-file gpurun_out/r2_linkwin_factor$f.csv python profiles/hot_step.py 2 > gpurun_out/ncu4.log 2>&1
echo "factor $f rc=$?"; grep link_window gpurun_out/r2_linkwin_factor$f.csv | tail -n 1 | awk -F'","' '{print $NF}'
done
timeout 600 ncu -k regex:link_window --set full --clock-control none --import-source on -s 1 -c 1 -o gpurun_out/r2_linkwin_full python profiles/hot_step.py 2 > gpurun_out/ncu4b.log 2>&1; echo "ncu full rc=$?"
