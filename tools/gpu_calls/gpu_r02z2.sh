#!/bin/bash
# round 2, call Z2 (1 GPU): ncu --set full of the N=2-shard default step kernel (kb_step: panel role + TMA pass role), final code
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
CMD="python tools/tune_blocked.py 10000 40000 3 --blocks 16 --mode 7 --variants -1"
timeout 60 $CMD > gpurun_out/r02z2_plain.log 2>&1 && \
timeout 100 ncu --set full --clock-control none --import-source on -k regex:kb_step -s 4 -c 1 -o /tmp/r02z2_step_n2 $CMD > gpurun_out/r02z2_ncu.log 2>&1
ncu -i /tmp/r02z2_step_n2.ncu-rep --page raw --csv > gpurun_out/r02z2_step_n2_raw.csv 2>/dev/null
tail -n 1 gpurun_out/r02z2_plain.log | cut -c1-300; tail -n 3 gpurun_out/r02z2_ncu.log; ls -la gpurun_out/r02z2_step_n2_raw.csv
