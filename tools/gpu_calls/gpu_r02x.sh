#!/bin/bash
# round 2, call X (1 GPU): the round-end checks on the final code: pytest -m gpu, smoke(), the default bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r02x_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02x_tests.log
tail -n 3 gpurun_out/r02x_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02x_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r02x_smoke.log
tail -n 2 gpurun_out/r02x_smoke.log
LPS_DEBUG=1 timeout 200 python bench.py > gpurun_out/r02x_bench_n1.json 2> gpurun_out/r02x_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r02x_bench_n1.err
cut -c1-300 gpurun_out/r02x_bench_n1.json; tail -n 2 gpurun_out/r02x_bench_n1.err | cut -c1-440
