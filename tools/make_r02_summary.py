"""Assemble profiles/r02_* from the files the round-2 evidence run (tools/gpu_r2_final.sh) left in gpurun_out/:
bench lines, phase timers (exact + variational), ncu launch list and --set full summaries.
usage: python tools/make_r02_summary.py"""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

def last_json(path):
    try:
        return json.loads(open(path).read().strip().splitlines()[-1])
    except Exception:
        return None

lines = []
bench = {}
for name in ("ref", "exact", "vartgp", "cfg5", "cfg2", "cfg3"):
    d = last_json(os.path.join(G, f"r2f_{name}.json"))
    if d is None:
        continue
    bench[name] = d
    tag = f"{d['value']:.0f}" if d["value"] >= 100 else f"{d['value']:.1f}"
    json.dump(d, open(os.path.join(P, f"r02_bench_{name}_{tag}.json"), "w"))
lines.append("# Round 2 -- measured on one B200 (tools/gpu_r2_final.sh; every file named below is in profiles/)\n")
lines.append("## Bench lines (`bench.py`, default flags unless stated)\n")
lines.append("| arm | value | e2e (product API, host buffers) | roofline frac | cpu_baseline (cores) | file |")
lines.append("|---|---|---|---|---|---|")
for name, d in bench.items():
    e = d.get("e2e") or {}
    r = d.get("roofline") or {}
    c = d.get("cpu_baseline") or {}
    tag = f"{d['value']:.0f}" if d["value"] >= 100 else f"{d['value']:.1f}"
    lines.append(f"| {name} | {d['value']:.1f} {d['unit']} ({d['ms_per_step']:.1f} ms/step, {d.get('gpu_launches')} launches) | "
                 f"{e.get('value', float('nan')):.1f} | {r.get('frac') if r else None} | {c.get('value')} ({c.get('cores')}) | r02_bench_{name}_{tag}.json |")
lines.append("")

# exact phase timers
d = None
try:
    d = json.load(open(os.path.join(G, "r2f_exact_phases.json")))
    json.dump(d, open(os.path.join(P, "r02_exact_phase_timers.json"), "w"), indent=1)
except Exception:
    pass
if d:
    lines.append("## k_exact_bo_step: cycles of thread 0 per closure (`tools/gpu_phase_timers.py`, 296 campaigns of n points, one BO iteration)\n")
    keys = ["gram", "sweep_A", "sweep_B", "sweep_C", "gradient", "reduce", "advance"]
    lines.append("| n | closures / fit | launch ms | " + " | ".join(keys) + " | pool pass share | factorise share |")
    lines.append("|---|---|---|" + "---|" * (len(keys) + 2))
    for k, v in d.items():
        lines.append(f"| {k} | {v['closures_per_fit']:.1f} | {v['launch_ms']:.2f} | " + " | ".join(f"{v['cycles_per_closure'][q]:.0f}" for q in keys) +
                     f" | {v['share']['pool_pass']:.3f} | {v['share']['factorise']:.3f} |")
    lines.append("")

# variational phase timers
vt = []
for f in ("r2f_var_phases.json", "r2f_var_phases_cfg5.json"):
    try:
        vt.append(json.load(open(os.path.join(G, f))))
    except Exception:
        pass
if vt:
    out = ["# var_epoch inside k_var_fit: cycles of thread 0 per epoch (`tools/gpu_phase_timers_var.py`, VarTGP)\n",
           "Round-1 build (column-by-column Cholesky-inverse, pair sweep above cap 87, unstaged contractions everywhere), call 13:",
           "n=60/B=296: 0.280 ms per epoch, Ct_Kzz_cholinv 135 366, contractions 198 259, quadrature 56 417, grad_pass 63 770;",
           "n=89/B=296: 0.684 ms, sweep_P 330 858 (pair sweep), cholinv 307 441; n=256/B=148: 18.86 ms, sweep_P 12 967 837, cholinv 14 987 453,",
           "contractions 5 892 460, grad_pass 1 019 893 (gpurun_out/r2c13_var_phases*.json).\n", "Current build:\n"]
    for d in vt:
        for k, v in d.items():
            c = v["cycles_per_epoch"]
            contr = sum(c[q] for q in c if q.startswith("c") and q[1].isdigit())
            sweeps = c.get("sweep_A", 0) + c.get("sweep_B", 0) + c.get("sweep_C", 0) + c.get("setup_sweep_P", 0)
            out.append(f"* {k}: {v['ms_per_epoch']:.3f} ms per epoch (launch of {v['launch_ms']:.1f} ms); sweeps (P and K_ZZ, incl. fills) {sweeps}, "
                       f"contractions {contr}, quadrature + loss {c['quadrature_loss']}, kernel-gradient pass {c['grad_pass']}, Adam {c['adam']}")
            out.append("  " + ", ".join(f"{a} {b}" for a, b in c.items()))
    open(os.path.join(P, "r02_var_phase_timers.md"), "w").write("\n".join(out) + "\n")
    lines.append("## Variational epoch: see r02_var_phase_timers.md, r02_var_ab.md, r02_dmma_probe.txt\n")

# ncu launch list
f = os.path.join(G, "r2f_launches.csv")
if os.path.exists(f):
    shutil.copy(f, os.path.join(P, "r02_launches_bench_steps12.csv"))
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    head = next((r for r in rows if "Kernel Name" in r), None)
    if head:
        ki, vi, bi, gi = head.index("Kernel Name"), head.index("Metric Value"), head.index("Block Size"), head.index("Grid Size")
        tot = {}
        for r in rows:
            if r is head:
                continue
            try:
                tot.setdefault((r[ki][:48], r[bi], r[gi]), []).append(float(r[vi].replace(",", "")))
            except ValueError:
                pass
        all_t = sum(sum(v) for v in tot.values()) or 1.0
        lines.append("## ncu launch list of `bench.py --steps 12 --warmup 3` (`-k regex:k_exact`, launches 160..; r02_launches_bench_steps12.csv)\n")
        s12, g1 = last_json(os.path.join(G, "r2f_exact_steps12.json")) or {}, last_json(os.path.join(G, "r2f_exact_groups1.json")) or {}
        lines.append("Only `k_exact_bo_step` runs inside the timed region (one launch per cohort and step: its share of the step is 100 %).  ncu serialises "
                     f"the launches: a step is 16 launches of 64 campaigns, 16 x ~8 ms one after another against {s12.get('ms_per_step', float('nan')):.1f} ms per step "
                     f"measured with the 16 streams overlapping (the same command without ncu: {s12.get('value', float('nan')):.0f} it/s at 12 steps; a single "
                     f"1024-campaign launch per step, `--groups 1 --schedule mixed`: {g1.get('ms_per_step', float('nan')):.1f} ms, {g1.get('value', float('nan')):.0f} it/s; "
                     "both lines are committed as r02_bench_exact_steps12_*.json / r02_bench_exact_groups1_*.json).\n")
        for k, v in sorted(tot.items(), key=lambda kv: -sum(kv[1])):
            lines.append(f"* {k[0]} block {k[1]} grid {k[2]}: {len(v)} launches, mean {sum(v) / len(v) / 1e6:.2f} ms, {100 * sum(v) / all_t:.1f} % of the listed kernel time")
        lines.append("")

# ncu --set full summaries (the k_exact_bo_step / k_var_fit captures are one build older than the bench lines above: taken
# before the K_ZZ retention and the quadrature reciprocals, which do not change the shape of those kernels)
lines.append("The `ncu --set full` captures of k_exact_bo_step / k_var_fit below are one build older than the bench lines (before the K_ZZ "
             "retention and the hoisted quadrature reciprocals); the k_var_acq captures are of the final build (both CTA shapes).\n")
for rep, name in (("prof_r2f_cohort_n60", "r02_ncu_final_cohort_n60"), ("prof_r2f_varfit_n60", "r02_ncu_final_varfit_n60"),
                  ("prof_r2f_varfit_cfg5", "r02_ncu_final_varfit_cfg5"), ("prof_r2f_varacq_n50", "r02_ncu_final_varacq_n50"),
                  ("prof_r2f_varacq_n89", "r02_ncu_final_varacq_n89")):
    src = os.path.join(G, rep + ".ncu-rep")
    if os.path.exists(src):
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), src], capture_output=True, text=True).stdout
        open(os.path.join(P, name + ".txt"), "w").write(txt)
        lines.append(f"## {name}.txt (`ncu --set full`, one launch)\n")
        keep = [l for l in txt.splitlines() if any(s in l for s in ("gpu__time_duration", "dram__bytes_read.sum [", "dram__bytes_write.sum [", "pipe_fp64_cycles_active",
                "issue_active", "warps_active", "registers_per_thread [", "shared_mem_per_block_dynamic", "subpipe_dmma", "pcsamp"))]
        lines += ["    " + l for l in keep[:16]] + [""]
open(os.path.join(P, "r02_summary.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
