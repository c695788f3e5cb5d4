#!/bin/bash
# round 2, call K (2 GPUs): panel role with two cells per thread + L2 eviction hints: parity, panel clock, N=2 bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export LPS_DEBUG=1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02k_all.log 2>&1
echo "all rc=$?" >> gpurun_out/r02k_all.log
timeout 600 python tools/tune_blocked.py 20000 40000 8 --blocks 16 --mode 7 --variants -1 --panel 0,4,6,8 >> gpurun_out/r02k_tune.jsonl 2>> gpurun_out/r02k_tune.err
timeout 600 python tools/tune_blocked.py 2500 40000 12 --blocks 16 --mode 7 --variants -1 --panel 0,8,12,16,24,32 >> gpurun_out/r02k_tune.jsonl 2>> gpurun_out/r02k_tune.err
timeout 600 python tools/tune_blocked.py 5000 40000 12 --blocks 16 --mode 7 --variants -1 --panel 0,8,12,16,24 >> gpurun_out/r02k_tune.jsonl 2>> gpurun_out/r02k_tune.err
timeout 600 python tools/tune_blocked.py 10000 40000 8 --blocks 16 --mode 7 --variants -1 --panel 0,6,8,12,16 >> gpurun_out/r02k_tune.jsonl 2>> gpurun_out/r02k_tune.err
timeout 600 python tools/tune_blocked.py 10000 10000 12 --blocks 16 --mode 7 --variants -1 --panel 0,4,8,16 >> gpurun_out/r02k_tune.jsonl 2>> gpurun_out/r02k_tune.err
run() {
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e $2 > gpurun_out/r02k_bench_$1.json 2> gpurun_out/r02k_bench_$1.err
  echo "rc=$?" >> gpurun_out/r02k_bench_$1.err
}
run n2_m7 "--loop-mode 7"
run n2_m7P6 "--loop-mode 7 --panel-ctas 6"
run n2_m7P12 "--loop-mode 7 --panel-ctas 12"
tail -n 4 gpurun_out/r02k_all.log
