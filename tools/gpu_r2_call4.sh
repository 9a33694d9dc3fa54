#!/bin/bash
mkdir -p gpurun_out
line() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), d['frozen_campaigns'], d['clocks']['sm_mhz'], round(d['e2e']['value']), d['gpu_launches'], d.get('parity'))"; }
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -o stpbenchmark_b200/libstp_b200_r1.so tools/ab_r1/csrc/stp_kernels.cu
echo "== r1 library, cohort"; for r in 1 2; do STP_LIB=$PWD/stpbenchmark_b200/libstp_b200_r1.so python bench.py --steps 40 --warmup 3 --no-cpu-baseline 2>/dev/null | line; done
echo "== new library, cohort"; for r in 1 2; do python bench.py --steps 40 --warmup 3 --no-cpu-baseline 2>/dev/null | line; done
echo "== new library, queue"; for r in 1 2; do python bench.py --steps 40 --warmup 3 --no-cpu-baseline --schedule queue 2>/dev/null | line; done
echo "== new library, queue, threads 128/192"; for t in 128 192; do STP_THREADS=$t python bench.py --steps 40 --warmup 3 --no-cpu-baseline --schedule queue 2>/dev/null | line; done
echo "== new, queue with cpu parity"; python bench.py --steps 40 --warmup 3 --schedule queue 2>/dev/null | tee gpurun_out/r2c4_bench_queue.json | line
echo "== phase timers"
python tools/gpu_phase_timers.py --n 30 60 89 --B 296 592 > gpurun_out/r2c4_phase_timers.json 2> gpurun_out/r2c4_phase_timers.err; tail -3 gpurun_out/r2c4_phase_timers.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c4_phase_timers.json'))
for k,v in d.items(): print(k, round(v['launch_ms'],2), v['cycles_per_closure'], v.get('advance_split'))
PY
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
