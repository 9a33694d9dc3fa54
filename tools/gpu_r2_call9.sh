#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name: $*"; timeout 1200 "$@" > gpurun_out/r2c9_$name.json 2> gpurun_out/r2c9_$name.err; echo "rc=$?"; tail -c 400 gpurun_out/r2c9_$name.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2c9_$name.json').read().strip().splitlines()[-1])
    print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('value','ms_per_step','steps','gpu_launches','frozen_campaigns','setup_s','designs_s','parity')}, 'frac', d.get('roofline') and round(d['roofline']['frac'],4), 'e2e', d['e2e'] and {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k!='what'}, 'cpu', d['cpu_baseline'] and {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['cpu_baseline'].items() if k in ('value','cores','per_core')})
except Exception as e: print('no line', e)
PY
}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run exact python bench.py
run ref python bench.py --impl reference --steps 40 --warmup 3
run vartgp python bench.py --model VarTGP --steps 3
run cfg2 python bench.py --config cfg2
run cfg5 python bench.py --config cfg5
for spec in "40 128" "60 160" "85 256"; do set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:k_exact_bo_step -s 1 -c 1 -f -o gpurun_out/prof_r2_cohort_n$1 python tools/gpu_single_n.py --n $1 --B 64 > gpurun_out/ncu_r2_cohort_n$1.log 2>&1; tail -1 gpurun_out/ncu_r2_cohort_n$1.log
done
ncu --set full --clock-control none --import-source on -k regex:k_var_fit -s 1 -c 1 -f -o gpurun_out/prof_r2_varfit_n60 python tools/gpu_prof_var.py 60 > gpurun_out/ncu_r2_varfit.log 2>&1; tail -1 gpurun_out/ncu_r2_varfit.log
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 40 -c 200 --csv --log-file gpurun_out/r2c9_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c9_ncu_launch.log 2>&1; tail -2 gpurun_out/r2c9_launches.csv | cut -c1-200
