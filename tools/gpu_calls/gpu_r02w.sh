#!/bin/bash
# round 2, call W (8 GPUs): the multi-GPU parity tests on 2 / 4 / 8 ranks and the driver's N=8 line, final code
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_blocked.py tests/test_gpu_sharded.py -m gpu -q -k "multi_gpu or sharded or shard or world or gpus" > gpurun_out/r02w_multi.log 2>&1
echo "multi rc=$?" >> gpurun_out/r02w_multi.log
export LPS_DEBUG=1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
  bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02w_bench_n8.json 2> gpurun_out/r02w_bench_n8.err
echo "rc=$?" >> gpurun_out/r02w_bench_n8.err
tail -n 3 gpurun_out/r02w_multi.log; cut -c1-200 gpurun_out/r02w_bench_n8.json; tail -n 1 gpurun_out/r02w_bench_n8.err
grep -h "rank 0  panel role" gpurun_out/r02w_bench_n8.err | tail -n 2 | cut -c1-440
