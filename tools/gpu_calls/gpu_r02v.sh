#!/bin/bash
# round 2, call V (8 GPUs): the N=8 and N=4 lines with the self-tuned split
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export LPS_DEBUG=1
run() {
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $2 --steps 20 --warmup 5 --no-e2e $3 > gpurun_out/r02v_bench_$1.json 2> gpurun_out/r02v_bench_$1.err
  echo "rc=$?" >> gpurun_out/r02v_bench_$1.err
}
run n8 8 ""
run n4 4 ""
grep -h "rank 0  panel role" gpurun_out/r02v_bench_n8.err | tail -n 3 | cut -c1-440
for f in gpurun_out/r02v_bench_*.json; do echo $f; cut -c1-200 $f; done
grep -h "rank 0  panel role" gpurun_out/r02v_bench_n4.err | tail -n 2 | cut -c1-440
