#!/bin/bash
# Round-2 evidence run: GPU tests, every bench arm, phase timers, ncu launch list + three --set full captures.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
run() { name=$1; shift; timeout 1500 "$@" > gpurun_out/r2f_$name.json 2> gpurun_out/r2f_$name.err; echo "== $name rc=$? : $*"; tail -c 300 gpurun_out/r2f_$name.json | cut -c1-300; echo; }
run ref python bench.py --impl reference --steps 40 --warmup 3
run exact python bench.py
run vartgp python bench.py --model VarTGP --steps 3
run cfg5 python bench.py --config cfg5
run cfg2 python bench.py --config cfg2
run cfg3 python bench.py --config cfg3
python tools/gpu_phase_timers.py > gpurun_out/r2f_exact_phases.json 2> gpurun_out/r2f_exact_phases.err; echo "exact phases rc=$?"
python tools/gpu_phase_timers_var.py > gpurun_out/r2f_var_phases.json 2> gpurun_out/r2f_var_phases.err; echo "var phases rc=$?"
python tools/gpu_phase_timers_var.py cfg5 > gpurun_out/r2f_var_phases_cfg5.json 2>> gpurun_out/r2f_var_phases.err; echo "var phases cfg5 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_exact --launch-skip 160 -c 200 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2f_ncu_launch.log 2>&1; tail -2 gpurun_out/r2f_launches.csv | cut -c1-200
python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2f_exact_steps12.json 2>/dev/null
python bench.py --groups 1 --schedule mixed --steps 12 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2f_exact_groups1.json 2>/dev/null
[ -n "$SKIP_NCU_FULL" ] && exit 0     # the three captures below are ~20 MB each: gpurun brings back at most 64 MiB per call
ncu --set full --clock-control none --import-source on -k regex:k_exact_bo_step -s 1 -c 1 -f -o gpurun_out/prof_r2f_cohort_n60 python tools/gpu_single_n.py --n 60 --B 64 > gpurun_out/r2f_ncu_cohort_n60.log 2>&1; tail -1 gpurun_out/r2f_ncu_cohort_n60.log
ncu --set full --clock-control none --import-source on -k regex:k_var_fit -s 1 -c 1 -f -o gpurun_out/prof_r2f_varfit_n60 python tools/gpu_prof_var.py 60 > gpurun_out/r2f_ncu_varfit_n60.log 2>&1; tail -1 gpurun_out/r2f_ncu_varfit_n60.log
ncu --set full --clock-control none --import-source on -k regex:k_var_fit -s 1 -c 1 -f -o gpurun_out/prof_r2f_varfit_cfg5 python tools/gpu_prof_var.py 256 cfg5 > gpurun_out/r2f_ncu_varfit_cfg5.log 2>&1; tail -1 gpurun_out/r2f_ncu_varfit_cfg5.log
