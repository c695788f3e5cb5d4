#!/bin/bash
# round 2, call M (1 GPU): parity, default bench line, launch list, ncu --set full of the look-ahead step at the four
# shard sizes (summaries exported to CSV on the box: the .ncu-rep files are too big to bring back)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export LPS_DEBUG=1
timeout 900 python -m pytest tests/test_gpu_blocked.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02m_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02m_tests.log
timeout 900 python bench.py > gpurun_out/r02m_bench_n1.json 2> gpurun_out/r02m_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r02m_bench_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02m_bench_ref.json 2> gpurun_out/r02m_bench_ref.err
timeout 600 python bench.py --variant 13 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r02m_bench_n1_tma.json 2> gpurun_out/r02m_bench_n1_tma.err
timeout 600 python bench.py --loop-mode 6 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r02m_bench_n1_m6.json 2> gpurun_out/r02m_bench_n1_m6.err
timeout 600 python bench.py --panel-ctas 8 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r02m_bench_n1_P8.json 2> gpurun_out/r02m_bench_n1_P8.err
unset LPS_DEBUG
CMD="python bench.py --steps 2 --warmup 1 --pivots-per-step 64 --no-e2e --no-cpu-baseline --no-secondary"
$CMD > gpurun_out/r02m_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02m_launches.csv $CMD > gpurun_out/r02m_ncu_l.log 2>&1
for sz in "20000 40000 n1" "10000 40000 n2" "5000 40000 n4" "2500 40000 n8" "10000 10000 c3"; do
  set -- $sz
  CMD="python tools/tune_blocked.py $1 $2 3 --blocks 16 --mode 7 --variants -1"
  $CMD > gpurun_out/r02m_plain_$3.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:kb_step_flush -s 4 -c 1 -o /tmp/r02m_step_$3 $CMD > gpurun_out/r02m_ncu_$3.log 2>&1
  ncu -i /tmp/r02m_step_$3.ncu-rep --page raw --csv > gpurun_out/r02m_step_$3_raw.csv 2>/dev/null
  if [ "$3" = "n1" ] || [ "$3" = "n8" ]; then ncu -i /tmp/r02m_step_$3.ncu-rep --page source --csv > gpurun_out/r02m_step_$3_source.csv 2>/dev/null; fi
done
CMD="python tools/tune_blocked.py 20000 40000 3 --blocks 1 --mode 1 --variants -1"
$CMD > gpurun_out/r02m_plain_upd.log 2>&1 && \
ncu --set full --clock-control none -k regex:k_update -s 10 -c 1 -o /tmp/r02m_update $CMD > gpurun_out/r02m_ncu_upd.log 2>&1
ncu -i /tmp/r02m_update.ncu-rep --page raw --csv > gpurun_out/r02m_update_raw.csv 2>/dev/null
tail -n 2 gpurun_out/r02m_tests.log; cut -c1-300 gpurun_out/r02m_bench_n1.json; tail -n 3 gpurun_out/r02m_bench_n1.err; du -sh gpurun_out
