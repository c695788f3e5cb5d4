#!/bin/bash
# round 2, call E: pass shapes R8 (8 rows per thread) and F12 (folded producer, 12 consumer warps)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_blocked.py -m gpu -x -q > gpurun_out/r02e_blocked.log 2>&1
echo "blocked rc=$?" >> gpurun_out/r02e_blocked.log
for v in 0 12 13 14; do
  timeout 300 python tools/tune_blocked.py 20000 40000 8 --blocks 16 --mode 6 --variants $v >> gpurun_out/r02e_tune.jsonl 2>> gpurun_out/r02e_tune.err
done
timeout 600 python tools/tune_blocked.py 20000 40000 8 --blocks 16 --mode 7 --variants 12,13,14 --panel 8 --chunk 0,120 >> gpurun_out/r02e_tune.jsonl 2>> gpurun_out/r02e_tune.err
timeout 300 python tools/tune_blocked.py 2500 40000 12 --blocks 16 --mode 7 --variants 12,14 --panel 16,24 >> gpurun_out/r02e_tune.jsonl 2>> gpurun_out/r02e_tune.err
CMD="python tools/tune_blocked.py 20000 40000 3 --blocks 16 --mode 6 --variants 14"
$CMD > gpurun_out/r02e_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:kb_sweep -s 3 -c 1 -o gpurun_out/r02e_sweep_f12 $CMD > gpurun_out/r02e_ncu.log 2>&1
tail -n 3 gpurun_out/r02e_blocked.log; cut -c1-330 gpurun_out/r02e_tune.jsonl
