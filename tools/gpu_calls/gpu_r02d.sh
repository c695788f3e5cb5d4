#!/bin/bash
# round 2, call D: restructured pass pipeline (chunk descriptors, hinted waits, rolled special path)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_blocked.py -m gpu -x -q > gpurun_out/r02d_blocked.log 2>&1
echo "blocked rc=$?" >> gpurun_out/r02d_blocked.log
for v in 0 10 11 12; do
  timeout 300 python tools/tune_blocked.py 20000 40000 8 --blocks 16 --mode 6 --variants $v >> gpurun_out/r02d_tune.jsonl 2>> gpurun_out/r02d_tune.err
done
timeout 600 python tools/tune_blocked.py 20000 40000 8 --blocks 16 --mode 7 --variants 10,12 --panel 8 --chunk 0,120 >> gpurun_out/r02d_tune.jsonl 2>> gpurun_out/r02d_tune.err
timeout 300 python tools/tune_blocked.py 2500 40000 12 --blocks 16 --mode 7 --variants 10,12 --panel 16,24 >> gpurun_out/r02d_tune.jsonl 2>> gpurun_out/r02d_tune.err
timeout 300 python tools/tune_blocked.py 20000 40000 8 --blocks 8 --mode 7 --variants 10 --panel 8 >> gpurun_out/r02d_tune.jsonl 2>> gpurun_out/r02d_tune.err
CMD="python tools/tune_blocked.py 20000 40000 3 --blocks 16 --mode 6 --variants 12"
$CMD > gpurun_out/r02d_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:kb_sweep -s 3 -c 1 -o gpurun_out/r02d_sweep_t8 $CMD > gpurun_out/r02d_ncu.log 2>&1
CMD="python tools/tune_blocked.py 20000 40000 3 --blocks 16 --mode 6 --variants 10"
$CMD > gpurun_out/r02d_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:kb_sweep -s 3 -c 1 -o gpurun_out/r02d_sweep_w12 $CMD > gpurun_out/r02d_ncu2.log 2>&1
tail -n 3 gpurun_out/r02d_blocked.log; cut -c1-330 gpurun_out/r02d_tune.jsonl
