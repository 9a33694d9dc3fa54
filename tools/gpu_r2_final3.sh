#!/bin/bash
# final lines of the round with 24 cohorts: driver flags, then default flags
python bench.py --steps 20 --warmup 5 > gpurun_out/r2y_exact_20.json 2> gpurun_out/r2y_exact_20.err
python bench.py > gpurun_out/r2y_exact.json 2> gpurun_out/r2y_exact.err
python - <<PY
import json
for f in ("r2y_exact_20", "r2y_exact"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4),
          "cpu", round(d["cpu_baseline"]["value"], 1), d["parity"], d["frozen_campaigns"], d["gpu_launches"], d["clocks"])
PY
