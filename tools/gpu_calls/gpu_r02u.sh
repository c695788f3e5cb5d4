#!/bin/bash
# round 2, call U (2 GPUs): sharded parity tests and the N=2 line with the self-tuned split
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export LPS_DEBUG=1
timeout 300 python -m pytest tests/test_gpu_blocked.py tests/test_gpu_sharded.py -m gpu -q -k "multi_gpu or sharded or shard or world or gpus" > gpurun_out/r02u_multi.log 2>&1
echo "multi rc=$?" >> gpurun_out/r02u_multi.log
run() {
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $2 --steps 20 --warmup 5 --no-e2e $3 > gpurun_out/r02u_bench_$1.json 2> gpurun_out/r02u_bench_$1.err
  echo "rc=$?" >> gpurun_out/r02u_bench_$1.err
}
run n2 2 ""
tail -n 3 gpurun_out/r02u_multi.log; for f in gpurun_out/r02u_bench_*.json; do echo $f; cut -c1-200 $f; done
grep -h "rank 0  panel role" gpurun_out/r02u_bench_n2.err | tail -n 3 | cut -c1-440
