#!/bin/bash
# round 2, call J (8 GPUs): multi-GPU parity tests, then the look-ahead loop at 8 / 4 / 2 / 1 ranks
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export LPS_DEBUG=1
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r02j_gpus.txt
timeout 900 python -m pytest tests -m gpu -q -k "multi_gpu or sharded or shard or world or two_gpus or gpus" > gpurun_out/r02j_multi.log 2>&1
echo "multi rc=$?" >> gpurun_out/r02j_multi.log
run() {  # name, nproc, extra args
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $2 --steps 8 --warmup 3 --no-e2e $3 > gpurun_out/r02j_bench_$1.json 2> gpurun_out/r02j_bench_$1.err
  echo "rc=$?" >> gpurun_out/r02j_bench_$1.err
}
run n8_m7 8 "--loop-mode 7"
run n8_m6 8 "--loop-mode 6"
run n8_m7P16 8 "--loop-mode 7 --panel-ctas 16"
run n8_m7P32 8 "--loop-mode 7 --panel-ctas 32"
run n8_m7P48 8 "--loop-mode 7 --panel-ctas 48"
run n8_m7tma 8 "--loop-mode 7 --variant 13"
run n4_m7 4 "--loop-mode 7"
run n4_m6 4 "--loop-mode 6"
run n2_m7 2 "--loop-mode 7"
for f in gpurun_out/r02j_bench_*.json; do echo $f; cut -c1-260 $f; done; tail -n 4 gpurun_out/r02j_multi.log; grep -h "panel role" gpurun_out/r02j_bench_n8_m7.err | tail -n 3
