#!/bin/bash
# round 2, call B: ncu --set full of the TMA pass (in place, mode 6) and of the fused look-ahead step (mode 7)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
CMD6="python tools/tune_blocked.py 20000 40000 3 --blocks 16 --mode 6 --variants 10"
CMD7="python tools/tune_blocked.py 20000 40000 3 --blocks 16 --mode 7 --variants 10 --panel 8"
$CMD6 > gpurun_out/r02b_plain6.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:kb_sweep -s 2 -c 2 -o gpurun_out/r02b_sweep $CMD6 > gpurun_out/r02b_ncu6.log 2>&1
$CMD7 > gpurun_out/r02b_plain7.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:kb_step -s 3 -c 2 -o gpurun_out/r02b_step $CMD7 > gpurun_out/r02b_ncu7.log 2>&1
tail -2 gpurun_out/r02b_plain6.log gpurun_out/r02b_plain7.log; tail -3 gpurun_out/r02b_ncu6.log gpurun_out/r02b_ncu7.log; ls -la gpurun_out/r02b_*
