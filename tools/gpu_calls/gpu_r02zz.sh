#!/bin/bash
# round 2, call ZZ (1 GPU): two sharded ranks on one device (ordinary-launch loop shapes)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 55 python -m pytest tests/test_gpu_sharded.py -m gpu -k share_one_gpu -q > gpurun_out/r02zz_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02zz_tests.log
tail -n 25 gpurun_out/r02zz_tests.log | cut -c1-200
