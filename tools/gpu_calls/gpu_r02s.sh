#!/bin/bash
# round 2, call S (1 GPU): bulk-copy staged panel trips (LPS_PANEL_CELLS=3) against the barrier-per-trip form (=2), split tuning on
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_blocked.py tests/test_gpu_generators.py -m gpu -x -q > gpurun_out/r02s_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02s_tests.log
tail -n 4 gpurun_out/r02s_tests.log
export LPS_DEBUG=1
b() { timeout 400 python bench.py --steps 10 --warmup 4 --no-e2e --no-cpu-baseline --no-secondary $2 > gpurun_out/r02s_bench_$1.json 2> gpurun_out/r02s_bench_$1.err; echo "rc=$?" >> gpurun_out/r02s_bench_$1.err; }
b n1_c3 ""
LPS_PANEL_CELLS=2 LPS_SPLIT_TUNE=0 b n1_c2 ""
for sz in "2500 40000 12" "5000 40000 12"; do
  set -- $sz
  for c in 3 2; do
    echo "# cells=$c $1x$2" >> gpurun_out/r02s_tune.jsonl; echo "# cells=$c $1x$2" >> gpurun_out/r02s_tune.err
    LPS_PANEL_CELLS=$c timeout 300 python tools/tune_blocked.py $1 $2 $3 --blocks 16 --mode 7 --variants -1 --panel 24,32,40,0 >> gpurun_out/r02s_tune.jsonl 2>> gpurun_out/r02s_tune.err
  done
done
for f in gpurun_out/r02s_bench_*.json; do echo $f; cut -c1-160 $f; done
grep -h "panel role" gpurun_out/r02s_bench_n1_c3.err | tail -n 2 | cut -c1-420
grep -h "panel role" gpurun_out/r02s_bench_n1_c2.err | tail -n 1 | cut -c1-420
cat gpurun_out/r02s_tune.jsonl | cut -c1-330
