#!/bin/bash
# what the driver runs at round end (GPU tests, smoke, reference arm, default bench arm with its flags) on the ABI v3 build
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2z_ref.json 2> gpurun_out/r2z_ref.err; cut -c1-200 gpurun_out/r2z_ref.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_exact_20.json 2> gpurun_out/r2z_exact_20.err; cut -c1-160 gpurun_out/r2z_exact_20.json
python bench.py > gpurun_out/r2z_exact.json 2> gpurun_out/r2z_exact.err; cut -c1-160 gpurun_out/r2z_exact.json
python - <<PY
import json
for f in ("r2z_exact_20", "r2z_exact"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4),
          "cpu", round(d["cpu_baseline"]["value"], 1), d["parity"], d["frozen_campaigns"], d["clocks"])
PY
