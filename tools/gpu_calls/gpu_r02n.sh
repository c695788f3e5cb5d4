#!/bin/bash
# round 2, call N (1 GPU): warp-specialised look-ahead step (mode 8): parity, shard timings, sustained bench lines
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export LPS_DEBUG=1
timeout 900 python -m pytest tests/test_gpu_blocked.py tests/test_gpu_generators.py -m gpu -x -q > gpurun_out/r02n_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02n_tests.log
for sz in "20000 40000 8" "10000 40000 8" "5000 40000 12" "2500 40000 12" "10000 10000 12"; do
  set -- $sz
  timeout 300 python tools/tune_blocked.py $1 $2 $3 --blocks 16 --mode 8 --variants -1 >> gpurun_out/r02n_tune.jsonl 2>> gpurun_out/r02n_tune.err
done
b() { timeout 600 python bench.py --no-e2e --no-cpu-baseline --no-secondary $2 > gpurun_out/r02n_bench_$1.json 2> gpurun_out/r02n_bench_$1.err; }
b m8 "--loop-mode 8"
b m7tmaP6 "--loop-mode 7 --variant 13 --panel-ctas 6"
b m7tmaP10 "--loop-mode 7 --variant 13 --panel-ctas 10"
b m7flushP6 "--loop-mode 7 --panel-ctas 6"
b m6tma "--loop-mode 6 --variant 13"
tail -n 2 gpurun_out/r02n_tests.log; cut -c1-300 gpurun_out/r02n_tune.jsonl
