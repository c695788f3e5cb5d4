#!/bin/bash
# round 2, call F: generalized pass shapes A (12 warps R2 P2), B (15 warps, 192-col strips), C (T8), D (T8 P2), E (T8 R6)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_blocked.py -m gpu -x -q > gpurun_out/r02f_blocked.log 2>&1
echo "blocked rc=$?" >> gpurun_out/r02f_blocked.log
for v in 0 10 11 12 13 14; do
  timeout 300 python tools/tune_blocked.py 20000 40000 8 --blocks 16 --mode 6 --variants $v >> gpurun_out/r02f_tune.jsonl 2>> gpurun_out/r02f_tune.err
done
timeout 600 python tools/tune_blocked.py 20000 40000 8 --blocks 16 --mode 7 --variants 10,11,12,13,14 --panel 8 --chunk 120 >> gpurun_out/r02f_tune.jsonl 2>> gpurun_out/r02f_tune.err
timeout 300 python tools/tune_blocked.py 2500 40000 12 --blocks 16 --mode 7 --variants 10,11,12,13 --panel 24 >> gpurun_out/r02f_tune.jsonl 2>> gpurun_out/r02f_tune.err
timeout 300 python tools/tune_blocked.py 20000 40000 8 --blocks 8 --mode 7 --variants 10 --panel 8 >> gpurun_out/r02f_tune.jsonl 2>> gpurun_out/r02f_tune.err
for v in 10 11; do
CMD="python tools/tune_blocked.py 20000 40000 3 --blocks 16 --mode 6 --variants $v"
$CMD > gpurun_out/r02f_plain$v.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:kb_sweep -s 3 -c 1 -o gpurun_out/r02f_sweep_v$v $CMD > gpurun_out/r02f_ncu$v.log 2>&1
done
tail -n 3 gpurun_out/r02f_blocked.log; cut -c1-330 gpurun_out/r02f_tune.jsonl
