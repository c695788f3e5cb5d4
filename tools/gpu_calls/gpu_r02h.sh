#!/bin/bash
# round 2, call H (2 GPUs): multi-GPU parity tests incl. the look-ahead loop, then the 2-rank bench (torchrun)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r02h_gpus.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02h_all.log 2>&1
echo "all rc=$?" >> gpurun_out/r02h_all.log
run() {  # name, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e $2 > gpurun_out/r02h_bench_n2_$1.json 2> gpurun_out/r02h_bench_n2_$1.err
  echo "rc=$?" >> gpurun_out/r02h_bench_n2_$1.err
}
run m7 "--loop-mode 7"
run m6 "--loop-mode 6 --variant 0"
run m7flush "--loop-mode 7 --variant 0"
run m7P8 "--loop-mode 7 --panel-ctas 8"
run m7P24 "--loop-mode 7 --panel-ctas 24"
run m7flushP8 "--loop-mode 7 --variant 0 --panel-ctas 8"
timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --loop-mode 7 --variant 0 > gpurun_out/r02h_bench_n1_m7flush.json 2> gpurun_out/r02h_bench_n1_m7flush.err
timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --loop-mode 7 > gpurun_out/r02h_bench_n1_m7.json 2> gpurun_out/r02h_bench_n1_m7.err
tail -n 4 gpurun_out/r02h_all.log; for f in gpurun_out/r02h_bench_n*.json; do echo $f; cut -c1-300 $f; done; tail -n 3 gpurun_out/r02h_bench_n2_m7.err
