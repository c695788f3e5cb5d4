#!/bin/bash
# round 2, call P (8 GPUs): defaults (pass role by shard size, modelled panel/pass split) at 8 / 4 / 2 / 1 ranks
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export LPS_DEBUG=1
timeout 600 python -m pytest tests -m gpu -q -k "multi_gpu or sharded or shard or world or gpus" > gpurun_out/r02p_multi.log 2>&1
echo "multi rc=$?" >> gpurun_out/r02p_multi.log
run() {
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $2 --steps 20 --warmup 5 $3 > gpurun_out/r02p_bench_$1.json 2> gpurun_out/r02p_bench_$1.err
  echo "rc=$?" >> gpurun_out/r02p_bench_$1.err
}
run n8 8 ""
run n8_P24 8 "--no-e2e --panel-ctas 24"
run n8_P36 8 "--no-e2e --panel-ctas 36"
run n4 4 "--no-e2e"
run n2 2 "--no-e2e"
timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r02p_bench_n1.json 2> gpurun_out/r02p_bench_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29917 tools/c5_run.py --long-cap 8192 --out gpurun_out/r02p_c5_8gpu.json > gpurun_out/r02p_c5.log 2>&1
tail -n 3 gpurun_out/r02p_multi.log; for f in gpurun_out/r02p_bench_*.json; do echo $f; cut -c1-200 $f; done; grep -h "rank 0  panel role" gpurun_out/r02p_bench_n8.err | tail -n 2
