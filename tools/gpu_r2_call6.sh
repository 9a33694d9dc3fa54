#!/bin/bash
mkdir -p gpurun_out
line() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), d['frozen_campaigns'], d['clocks']['sm_mhz'], round(d['e2e']['value']), d['gpu_launches'], d.get('parity'))"; }
python -m pytest tests/test_gpu_exact.py tests/test_gpu_determinism.py -m gpu -x -q 2>&1 | tail -3
echo "== cohort"; for r in 1 2; do python bench.py --steps 40 --warmup 3 --no-cpu-baseline 2>/dev/null | line; done
echo "== cohort with parity"; python bench.py --steps 40 --warmup 3 2>/dev/null | tee gpurun_out/r2c6_bench.json | line
echo "== queue"; python bench.py --steps 40 --warmup 3 --no-cpu-baseline --schedule queue 2>/dev/null | line
echo "== queue trials=38 pad 0/28K/66K (4/2/1 per SM)"; for pad in 0 28672 67584; do STP_SMEM_PAD=$pad python bench.py --trials 38 --steps 40 --warmup 3 --no-cpu-baseline --schedule queue 2>/dev/null | line; done
echo "== queue trials=38 threads 64 / 96"; for t in 64 96; do STP_THREADS=$t python bench.py --trials 38 --steps 40 --warmup 3 --no-cpu-baseline --schedule queue 2>/dev/null | line; done
echo "== cohort threads 128 / 160 / 192"; for t in 128 160 192; do STP_THREADS=$t python bench.py --steps 40 --warmup 3 --no-cpu-baseline 2>/dev/null | line; done
echo "== phase timers"
python tools/gpu_phase_timers.py --n 30 60 89 --B 296 > gpurun_out/r2c6_phase_timers.json 2> gpurun_out/r2c6_phase_timers.err; tail -3 gpurun_out/r2c6_phase_timers.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c6_phase_timers.json'))
for k,v in d.items(): print(k, round(v['launch_ms'],2), v['cycles_per_closure'], v.get('advance_split'))
PY
