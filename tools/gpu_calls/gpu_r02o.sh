#!/bin/bash
# round 2, call O (2 GPUs): which pass role for the 2-rank shards, sustained (12 steps); N=1 TMA pass with fewer panel CTAs
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export LPS_DEBUG=1
run() {
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus 2 --steps 12 --warmup 3 --no-e2e $2 > gpurun_out/r02o_bench_$1.json 2> gpurun_out/r02o_bench_$1.err
  echo "rc=$?" >> gpurun_out/r02o_bench_$1.err
}
run n2_flush "--loop-mode 7"
run n2_tmaP8 "--loop-mode 7 --variant 13 --panel-ctas 8"
run n2_tmaP12 "--loop-mode 7 --variant 13 --panel-ctas 12"
run n2_tmaP16 "--loop-mode 7 --variant 13 --panel-ctas 16"
run n2_m6 "--loop-mode 6"
b() { timeout 600 python bench.py --steps 12 --no-e2e --no-cpu-baseline --no-secondary $2 > gpurun_out/r02o_bench_$1.json 2> gpurun_out/r02o_bench_$1.err; }
b n1_tmaP4 "--loop-mode 7 --variant 13 --panel-ctas 4"
b n1_tmaP5 "--loop-mode 7 --variant 13 --panel-ctas 5"
b n1_tmaP6 "--loop-mode 7 --variant 13 --panel-ctas 6"
for sz in "5000 40000 12" "2500 40000 12"; do
  set -- $sz
  timeout 300 python tools/tune_blocked.py $1 $2 $3 --blocks 16 --mode 7 --variants 13 --panel 16,24,32 >> gpurun_out/r02o_tune.jsonl 2>> gpurun_out/r02o_tune.err
  timeout 300 python tools/tune_blocked.py $1 $2 $3 --blocks 16 --mode 7 --variants -1 --panel 16,24,32 >> gpurun_out/r02o_tune.jsonl 2>> gpurun_out/r02o_tune.err
done
ls gpurun_out/r02o_*
