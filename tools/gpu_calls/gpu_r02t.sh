#!/bin/bash
# round 2, call T (1 GPU): pending rows persisting in L2 (LPS_L2_PERSIST) on / off, both staging forms
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_blocked.py -m gpu -x -q > gpurun_out/r02t_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02t_tests.log
tail -n 3 gpurun_out/r02t_tests.log
export LPS_DEBUG=1
b() { timeout 400 python bench.py --steps 10 --warmup 4 --no-e2e --no-cpu-baseline --no-secondary $2 > gpurun_out/r02t_bench_$1.json 2> gpurun_out/r02t_bench_$1.err; echo "rc=$?" >> gpurun_out/r02t_bench_$1.err; }
LPS_PANEL_CELLS=2 b n1_c2_pin ""
LPS_PANEL_CELLS=2 LPS_L2_PERSIST=0 b n1_c2_nopin ""
LPS_PANEL_CELLS=3 b n1_c3_pin ""
for c in 2 3; do for pin in 1 0; do
  echo "# cells=$c pin=$pin 2500x40000" >> gpurun_out/r02t_tune.jsonl; echo "# cells=$c pin=$pin" >> gpurun_out/r02t_tune.err
  LPS_SPLIT_TUNE=0 LPS_L2_PERSIST=$pin LPS_PANEL_CELLS=$c timeout 300 python tools/tune_blocked.py 2500 40000 12 --blocks 16 --mode 7 --variants -1 --panel 16,24,32 >> gpurun_out/r02t_tune.jsonl 2>> gpurun_out/r02t_tune.err
done; done
for f in gpurun_out/r02t_bench_*.json; do echo $f; cut -c1-100 $f; done
for f in gpurun_out/r02t_bench_*.err; do echo $f; grep -h "pinned" $f | head -n 1; grep -h "panel role" $f | tail -n 1 | cut -c30-420; done
sed -E 's/"chunk_rows.*"pivots_per_s": ([0-9.]+).*"pass_ms": ([0-9.]+).*/ pps=\1 pass_ms=\2/' gpurun_out/r02t_tune.jsonl | cut -c1-160
grep -h "panel role" gpurun_out/r02t_tune.err | cut -c30-330
