#!/bin/bash
# round 2, GPU call 2: persistent work-queue kernel vs cohort launches; hardware-queue count; occupancy with unlimited supply
mkdir -p gpurun_out
line() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), d['frozen_campaigns'], d['clocks']['sm_mhz'], round(d['e2e']['value']), d['gpu_launches'])"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== cohort default"
python bench.py --steps 40 --warmup 3 --no-cpu-baseline 2>/dev/null | line
echo "== cohort, CUDA_DEVICE_MAX_CONNECTIONS=32, groups 16 / 32 / 64"
for g in 16 32 64; do
  CUDA_DEVICE_MAX_CONNECTIONS=32 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --groups $g 2>/dev/null | line
done
echo "== queue (persistent), K=40"
python bench.py --steps 40 --warmup 3 --schedule queue > gpurun_out/r2c2_bench_queue.json 2> gpurun_out/r2c2_bench_queue.err; tail -2 gpurun_out/r2c2_bench_queue.err; line < gpurun_out/r2c2_bench_queue.json
python bench.py --steps 40 --warmup 3 --no-cpu-baseline --schedule queue 2>/dev/null | line
echo "== queue, threads 128 / 192"
for thr in 128 192; do
  STP_THREADS=$thr python bench.py --steps 40 --warmup 3 --no-cpu-baseline --schedule queue 2>/dev/null | line
done
echo "== queue, trials=38 (n<=47,128 thr) pad 0 / 28K / 66K -> 4 / 3 / 2 per SM"
for pad in 0 28672 67584; do
  STP_SMEM_PAD=$pad python bench.py --trials 38 --steps 40 --warmup 3 --no-cpu-baseline --schedule queue 2>/dev/null | line
done
echo "== queue, trials=38, threads 64 / 256"
for thr in 64 256; do
  STP_THREADS=$thr python bench.py --trials 38 --steps 40 --warmup 3 --no-cpu-baseline --schedule queue 2>/dev/null | line
done
echo "== phase timers"
python tools/gpu_phase_timers.py --n 30 60 89 --B 148 296 592 > gpurun_out/r2c2_phase_timers.json 2> gpurun_out/r2c2_phase_timers.err; tail -3 gpurun_out/r2c2_phase_timers.err; cat gpurun_out/r2c2_phase_timers.json
