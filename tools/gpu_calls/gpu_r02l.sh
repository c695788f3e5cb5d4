#!/bin/bash
# round 2, call L (2 GPUs): which L2 eviction hint is the illegal instruction; then parity + timing with the legal ones
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export LPS_DEBUG=1
for hm in 0 1 2 4; do
  LPS_L2_HINTS=$hm timeout 120 python -c "
import numpy as np, linear_programming_solver_b200 as L
from oracle import tier_f
A,b,c = tier_f.gen_dense_feasible(300,500,3)
ref = tier_f.TierFState(A.copy(),b.copy(),c.copy()); ref.run()
st = L.LPState(A,b,c,300,500,loop_mode=7); r = st.run()
print('hints $hm ok', r.npivots, st.pivot_log == ref.log, np.array_equal(st.b, ref.b))
" > gpurun_out/r02l_hint$hm.log 2>&1
  echo "hints=$hm rc=$?" >> gpurun_out/r02l_hints.log; tail -n 1 gpurun_out/r02l_hint$hm.log >> gpurun_out/r02l_hints.log
done
LPS_L2_HINTS=0 timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02l_all.log 2>&1
echo "all rc=$?" >> gpurun_out/r02l_all.log
for hm in 0 1 2 4; do
if grep -q "hints $hm ok" gpurun_out/r02l_hint$hm.log; then
LPS_L2_HINTS=$hm timeout 300 python tools/tune_blocked.py 2500 40000 12 --blocks 16 --mode 7 --variants -1 --panel 16,24 >> gpurun_out/r02l_tune_h$hm.jsonl 2>> gpurun_out/r02l_tune_h$hm.err
LPS_L2_HINTS=$hm timeout 300 python tools/tune_blocked.py 20000 40000 8 --blocks 16 --mode 7 --variants -1 --panel 6,8 >> gpurun_out/r02l_tune_h$hm.jsonl 2>> gpurun_out/r02l_tune_h$hm.err
fi
done
export LPS_L2_HINTS=0
timeout 600 python tools/tune_blocked.py 2500 40000 12 --blocks 16 --mode 7 --variants -1 --panel 0,8,12,32 >> gpurun_out/r02l_tune.jsonl 2>> gpurun_out/r02l_tune.err
timeout 600 python tools/tune_blocked.py 5000 40000 12 --blocks 16 --mode 7 --variants -1 --panel 0,8,12,16,24 >> gpurun_out/r02l_tune.jsonl 2>> gpurun_out/r02l_tune.err
timeout 600 python tools/tune_blocked.py 10000 40000 8 --blocks 16 --mode 7 --variants -1 --panel 0,6,8,12,16 >> gpurun_out/r02l_tune.jsonl 2>> gpurun_out/r02l_tune.err
timeout 600 python tools/tune_blocked.py 20000 40000 8 --blocks 16 --mode 7 --variants -1 --panel 0,4 >> gpurun_out/r02l_tune.jsonl 2>> gpurun_out/r02l_tune.err
timeout 600 python tools/tune_blocked.py 10000 10000 12 --blocks 16 --mode 7 --variants -1 --panel 0,4,8,16 >> gpurun_out/r02l_tune.jsonl 2>> gpurun_out/r02l_tune.err
run() {
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e $2 > gpurun_out/r02l_bench_$1.json 2> gpurun_out/r02l_bench_$1.err
  echo "rc=$?" >> gpurun_out/r02l_bench_$1.err
}
run n2_m7 "--loop-mode 7"
run n2_m7P12 "--loop-mode 7 --panel-ctas 12"
cat gpurun_out/r02l_hints.log; tail -n 4 gpurun_out/r02l_all.log
