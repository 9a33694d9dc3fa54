#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2x_exact_20.json 2> gpurun_out/r2x_exact_20.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2x_exact_20.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), d["e2e"].get("wall_s"), "frac", round(d["roofline"]["frac"], 4),
      "cpu", round(d["cpu_baseline"]["value"], 1), d["parity"], d["frozen_campaigns"], d["clocks"])
PY
