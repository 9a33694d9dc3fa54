#!/bin/bash
# round 2, GPU call 1: baseline, occupancy sensitivity (shared-memory padding at fixed thread count), phase timers
mkdir -p gpurun_out
line() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), d['frozen_campaigns'], d['clocks']['sm_mhz'], round(d['e2e']['value']))"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== baseline (default flags)"
python bench.py > gpurun_out/r2c1_bench_base.json 2> gpurun_out/r2c1_bench_base.err; line < gpurun_out/r2c1_bench_base.json
echo "== trials=38 (n<=47, 128 threads): pad 0 / 28K / 66K  -> 4 / 3 / 2 campaigns per SM"
for pad in 0 28672 67584; do
  for rep in 1 2; do
    STP_SMEM_PAD=$pad python bench.py --trials 38 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | line
  done
done
echo "== trials=56 (n<=65, 160 threads): pad 0 / 40K -> 3 / 2 per SM"
for pad in 0 40960; do
  STP_SMEM_PAD=$pad python bench.py --trials 56 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | line
done
echo "== full workload, threads 128 / 192 / auto"
for thr in 128 192 ""; do
  STP_THREADS=$thr python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | line
done
echo "== groups=1 mixed"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --groups 1 --schedule mixed 2>/dev/null | tee gpurun_out/r2c1_bench_groups1.json | line
echo "== phase timers"
python tools/gpu_phase_timers.py --n 20 30 45 60 75 89 > gpurun_out/r2c1_phase_timers.json 2> gpurun_out/r2c1_phase_timers.err; tail -3 gpurun_out/r2c1_phase_timers.err; cat gpurun_out/r2c1_phase_timers.json
