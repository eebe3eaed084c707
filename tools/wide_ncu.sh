#!/bin/bash
# ncu --set full of the eight wide-row kernel forms at 80 dims, condensed on the box (gpurun brings back at most 64 MiB).
# usage (on the GPU box, from the repo root): bash tools/wide_ncu.sh <tag>
T=${1:-wide}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spz_.*code_wide -f -o gpurun_out/${T}_ncu_wide80 \
    python tools/wide_probe.py > gpurun_out/${T}_ncu_wide80.log 2>&1
python tools/ncu_profile_csv.py gpurun_out/${T}_ncu_wide80_summary.csv gpurun_out/${T}_wide80_traffic.json wide80=gpurun_out/${T}_ncu_wide80.ncu-rep > /dev/null 2>&1
rm -f gpurun_out/${T}_ncu_wide80.ncu-rep
ls -la gpurun_out | tail -4
