#!/bin/bash
# A/B: number of age cohorts (streams) of the device-resident arm
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), round(d['ms_per_step'],2), d['gpu_launches'], d['frozen_campaigns'])"; }
for rep in 1 2; do
for g in ${GROUPS_LIST:-16 24 32}; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --groups $g 2>/dev/null | show "k20 groups=$g"
done
done
python bench.py --steps 20 --warmup 5 --groups 32 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('full k20 groups=32', round(d['value']), 'e2e', round(d['e2e']['value']), d['parity'], d['clocks'])"
