#!/bin/bash
# round 2, call I: out-of-line special path (parity + timing), panel-role clock vs number of panel CTAs
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export LPS_DEBUG=1
timeout 900 python -m pytest tests/test_gpu_blocked.py -m gpu -x -q > gpurun_out/r02i_blocked.log 2>&1
echo "blocked rc=$?" >> gpurun_out/r02i_blocked.log
for v in 12 13 14; do
  timeout 300 python tools/tune_blocked.py 20000 40000 8 --blocks 16 --mode 6 --variants $v >> gpurun_out/r02i_tune.jsonl 2>> gpurun_out/r02i_tune.err
done
timeout 600 python tools/tune_blocked.py 20000 40000 8 --blocks 16 --mode 7 --variants 0 --panel 1,2,3,4,6,8 >> gpurun_out/r02i_tune.jsonl 2>> gpurun_out/r02i_tune.err
timeout 600 python tools/tune_blocked.py 20000 40000 8 --blocks 16 --mode 7 --variants 13 --panel 2,4,8 --chunk 120 >> gpurun_out/r02i_tune.jsonl 2>> gpurun_out/r02i_tune.err
timeout 600 python tools/tune_blocked.py 2500 40000 12 --blocks 16 --mode 7 --variants 0 --panel 4,8,12,16,24,32 >> gpurun_out/r02i_tune.jsonl 2>> gpurun_out/r02i_tune.err
timeout 600 python tools/tune_blocked.py 10000 10000 12 --blocks 16 --mode 7 --variants 0 --panel 2,4,8,16 >> gpurun_out/r02i_tune.jsonl 2>> gpurun_out/r02i_tune.err
timeout 600 python tools/tune_blocked.py 10000 10000 12 --blocks 16 --mode 6 --variants 0 >> gpurun_out/r02i_tune.jsonl 2>> gpurun_out/r02i_tune.err
tail -n 3 gpurun_out/r02i_blocked.log; cut -c1-330 gpurun_out/r02i_tune.jsonl; grep "panel role" gpurun_out/r02i_tune.err | awk 'NR%2==0' | cut -c1-200
