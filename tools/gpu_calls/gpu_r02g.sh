#!/bin/bash
# round 2, call G: software-pipelined consumer (variant 14) vs shapes C (12) / D (13) and kb_flush (0)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_blocked.py -m gpu -x -q > gpurun_out/r02g_blocked.log 2>&1
echo "blocked rc=$?" >> gpurun_out/r02g_blocked.log
for v in 0 12 13 14; do
  timeout 300 python tools/tune_blocked.py 20000 40000 8 --blocks 16 --mode 6 --variants $v >> gpurun_out/r02g_tune.jsonl 2>> gpurun_out/r02g_tune.err
done
timeout 600 python tools/tune_blocked.py 20000 40000 8 --blocks 16 --mode 7 --variants 13,14 --panel 8 --chunk 120,240 >> gpurun_out/r02g_tune.jsonl 2>> gpurun_out/r02g_tune.err
timeout 300 python tools/tune_blocked.py 2500 40000 12 --blocks 16 --mode 7 --variants 13,14 --panel 24 >> gpurun_out/r02g_tune.jsonl 2>> gpurun_out/r02g_tune.err
CMD="python tools/tune_blocked.py 20000 40000 3 --blocks 16 --mode 6 --variants 14"
$CMD > gpurun_out/r02g_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:kb_sweep -s 3 -c 1 -o gpurun_out/r02g_sweep_v14 $CMD > gpurun_out/r02g_ncu.log 2>&1
tail -n 3 gpurun_out/r02g_blocked.log; cut -c1-330 gpurun_out/r02g_tune.jsonl
