#!/bin/bash
mkdir -p gpurun_out
for c in 5 10 1; do
python bench.py --chunk $c --no-cpu-baseline --no-e2e > gpurun_out/r2c22_chunk$c.json 2> gpurun_out/r2c22_chunk$c.err; echo "rc=$?"; tail -3 gpurun_out/r2c22_chunk$c.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c22_chunk$c.json').read().strip().splitlines()[-1])
print('chunk $c', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('value','ms_per_step','steps','warmup','gpu_launches','frozen_campaigns')}, 'frac', d['roofline']['frac'])
PY
done
python bench.py --chunk 5 --groups 8 --no-cpu-baseline --no-e2e > gpurun_out/r2c22_chunk5_g8.json 2>/dev/null; cut -c1-120 gpurun_out/r2c22_chunk5_g8.json
python bench.py --chunk 5 --steps 80 --no-cpu-baseline > gpurun_out/r2c22_chunk5_s80.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/r2c22_chunk5_s80.json').read().strip().splitlines()[-1])
print('chunk 5 s80', d['value'], d['e2e'])
PY
