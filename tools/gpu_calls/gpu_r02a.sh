#!/bin/bash
# round 2, call A: parity of the TMA pass + look-ahead loop, then a first timing sweep (one B200)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r02a_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_blocked.py -m gpu -x -q > gpurun_out/r02a_blocked.log 2>&1
echo "blocked rc=$?" >> gpurun_out/r02a_blocked.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02a_all.log 2>&1
echo "all rc=$?" >> gpurun_out/r02a_all.log
for cfg in "6 0" "6 10" "6 11" "6 12"; do
  set -- $cfg
  timeout 300 python tools/tune_blocked.py 20000 40000 8 --blocks 16 --mode $1 --variants $2 >> gpurun_out/r02a_tune.jsonl 2>> gpurun_out/r02a_tune.err
done
timeout 600 python tools/tune_blocked.py 20000 40000 8 --blocks 16 --mode 7 --variants 10,11,12 --panel 4,8,16 >> gpurun_out/r02a_tune.jsonl 2>> gpurun_out/r02a_tune.err
timeout 300 python tools/tune_blocked.py 20000 40000 8 --blocks 16 --mode 7 --variants 10 --panel 8 --chunk 60,120,480 >> gpurun_out/r02a_tune.jsonl 2>> gpurun_out/r02a_tune.err
# a small shard (what one of 8 ranks holds) and the 10,000 x 10,000 LP
timeout 300 python tools/tune_blocked.py 2500 40000 12 --blocks 16 --mode 6 --variants 0,10 >> gpurun_out/r02a_tune.jsonl 2>> gpurun_out/r02a_tune.err
timeout 300 python tools/tune_blocked.py 2500 40000 12 --blocks 16 --mode 7 --variants 10,11 --panel 8,16,32 >> gpurun_out/r02a_tune.jsonl 2>> gpurun_out/r02a_tune.err
timeout 300 python tools/tune_blocked.py 10000 10000 12 --blocks 16 --mode 7 --variants 10 --panel 8,16 >> gpurun_out/r02a_tune.jsonl 2>> gpurun_out/r02a_tune.err
tail -3 gpurun_out/r02a_blocked.log; tail -3 gpurun_out/r02a_all.log; cat gpurun_out/r02a_tune.jsonl | cut -c1-330
