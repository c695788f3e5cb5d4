#!/bin/bash
# round 2, call Y (1 GPU): re-check after the split planner became a pair of exported host functions
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/r02y_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02y_tests.log
tail -n 3 gpurun_out/r02y_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02y_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r02y_smoke.log
tail -n 2 gpurun_out/r02y_smoke.log
LPS_DEBUG=1 timeout 100 python bench.py --steps 8 --no-secondary --no-cpu-baseline --no-e2e > gpurun_out/r02y_bench_n1.json 2> gpurun_out/r02y_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r02y_bench_n1.err
cut -c1-200 gpurun_out/r02y_bench_n1.json; tail -n 2 gpurun_out/r02y_bench_n1.err | cut -c1-300
