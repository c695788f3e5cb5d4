#!/bin/bash
# round 2, call C: 4-columns-per-thread pass shapes, digests of the bench workload, first full bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_blocked.py -m gpu -x -q > gpurun_out/r02c_blocked.log 2>&1
echo "blocked rc=$?" >> gpurun_out/r02c_blocked.log
for v in 0 10 11 12; do
  timeout 300 python tools/tune_blocked.py 20000 40000 8 --blocks 16 --mode 6 --variants $v >> gpurun_out/r02c_tune.jsonl 2>> gpurun_out/r02c_tune.err
done
timeout 600 python tools/tune_blocked.py 20000 40000 8 --blocks 16 --mode 7 --variants 10,11,12 --panel 8 >> gpurun_out/r02c_tune.jsonl 2>> gpurun_out/r02c_tune.err
timeout 600 python tools/tune_blocked.py 20000 40000 8 --blocks 16 --mode 7 --variants 11 --panel 4,6,12 --chunk 0,120 >> gpurun_out/r02c_tune.jsonl 2>> gpurun_out/r02c_tune.err
timeout 300 python tools/tune_blocked.py 2500 40000 12 --blocks 16 --mode 7 --variants 10,11 --panel 16,24,32 >> gpurun_out/r02c_tune.jsonl 2>> gpurun_out/r02c_tune.err
timeout 300 python tools/tune_blocked.py 10000 10000 12 --blocks 16 --mode 7 --variants 10,11 --panel 8,16 >> gpurun_out/r02c_tune.jsonl 2>> gpurun_out/r02c_tune.err
timeout 300 python tools/tune_blocked.py 20000 40000 8 --blocks 8 --mode 7 --variants 10 --panel 8 >> gpurun_out/r02c_tune.jsonl 2>> gpurun_out/r02c_tune.err
timeout 600 python tools/make_bench_digests.py gpu 8192 gpurun_out/bench_digests_gpu.json > gpurun_out/r02c_digests.log 2>&1
timeout 900 python bench.py --variant 11 > gpurun_out/r02c_bench_v11.json 2> gpurun_out/r02c_bench_v11.err
echo "bench rc=$?" >> gpurun_out/r02c_bench_v11.err
timeout 600 python bench.py --loop-mode 6 --variant 0 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r02c_bench_old.json 2> gpurun_out/r02c_bench_old.err
CMD="python tools/tune_blocked.py 20000 40000 3 --blocks 16 --mode 6 --variants 11"
$CMD > gpurun_out/r02c_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:kb_sweep -s 3 -c 1 -o gpurun_out/r02c_sweep_w8 $CMD > gpurun_out/r02c_ncu.log 2>&1
tail -n 3 gpurun_out/r02c_blocked.log; cut -c1-330 gpurun_out/r02c_tune.jsonl; cat gpurun_out/r02c_digests.log; tail -n 5 gpurun_out/r02c_bench_v11.err; cut -c1-1500 gpurun_out/r02c_bench_v11.json; cut -c1-900 gpurun_out/r02c_bench_old.json
