#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name: $*"; timeout 900 "$@" > gpurun_out/r2c8_$name.json 2> gpurun_out/r2c8_$name.err; echo "rc=$?"; tail -c 600 gpurun_out/r2c8_$name.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2c8_$name.json').read().strip().splitlines()[-1])
    print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('value','ms_per_step','steps','gpu_launches','frozen_campaigns','setup_s','designs_s','parity')}, 'frac', d['roofline'] and round(d['roofline']['frac'],4), 'e2e', d['e2e'] and {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k!='what'}, 'cpu', d['cpu_baseline'] and {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['cpu_baseline'].items() if k in ('value','cores','per_core')}, d.get('per_model'))
except Exception as e: print('no line', e)
PY
}
run exact python bench.py
run ref python bench.py --impl reference --steps 8 --warmup 2
run vartgp python bench.py --model VarTGP --steps 3
run cfg2 python bench.py --config cfg2 --trials 20
run cfg3 python bench.py --config cfg3
run cfg5 python bench.py --config cfg5 --campaigns 148 --steps 1 --warmup 1
