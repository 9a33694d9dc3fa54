#!/bin/bash
mkdir -p gpurun_out
line() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), d['frozen_campaigns'], d['clocks']['sm_mhz'], round(d['e2e']['value']), d['gpu_launches'], d.get('parity'))"; }
python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -12
echo "== cohort"; python bench.py --steps 40 --warmup 3 --no-cpu-baseline 2>/dev/null | line
echo "== phase timers"
python tools/gpu_phase_timers.py --n 30 60 89 --B 296 > gpurun_out/r2c7_phase_timers.json 2> gpurun_out/r2c7_phase_timers.err; tail -3 gpurun_out/r2c7_phase_timers.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c7_phase_timers.json'))
for k,v in d.items(): print(k, round(v['launch_ms'],2), v['cycles_per_closure'], v.get('advance_split'))
PY
