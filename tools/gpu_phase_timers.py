"""Per-phase cycle budget of k_exact_bo_step ("first instrument").

Builds a -DSTP_PHASE_TIMERS variant of the library beside the shipped one, runs ONE BO iteration of B campaigns
that all have n training points (cfg #4 pools, Sobol design + random pool points) and prints, per n, the cycles
thread 0 of a CTA spent in every phase per closure / per campaign.  Run under gpurun:
    python tools/gpu_phase_timers.py [--n 30 60 89] [--B 296] > gpurun_out/phase_timers.json
"""
import argparse, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANT = os.path.join(ROOT, "stpbenchmark_b200", "libstp_b200_timers.so")


def build_variant():
    from stpbenchmark_b200 import build as b
    cmd = [b.nvcc_path()] + b.NVCC_FLAGS + ["-DSTP_PHASE_TIMERS", "-o", VARIANT] + b.SOURCES
    subprocess.run(cmd, check=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[30, 60, 89])
    ap.add_argument("--B", type=int, nargs="+", default=[296])
    args = ap.parse_args()
    build_variant()
    os.environ["STP_LIB"] = VARIANT
    import torch
    from stpbenchmark_b200 import _lib
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from gpu_single_n import build_batch
    dev = torch.device("cuda", 0)
    out = {}
    for n, B in [(n, B) for n in args.n for B in args.B]:
        batch = build_batch(n, B, dev)
        batch.step(); torch.cuda.synchronize()          # warm-up (a step changes n: rebuild)
        batch = build_batch(n, B, dev)
        _lib.phase_timers(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); batch.step(); e1.record(); torch.cuda.synchronize()
        t = _lib.phase_timers()
        F, C = max(t["closures"], 1), max(t["campaigns"], 1)
        per_closure = {k: round(t[k] / F, 1) for k in ("gram", "sweep_A", "sweep_B", "sweep_C", "gradient", "reduce", "advance")}
        per_campaign = {k: round(t[k] / C, 1) for k in ("load", "factorise", "pool_pass")}
        adv = {k: round(t[k] / F, 1) for k in ("adv_entry", "adv_copy_dot", "adv_dcsrch", "adv_projgr_tests", "adv_ls_step", "adv_build_B", "adv_cauchy", "adv_subsm", "adv_ls_start", "advance")}
        t["advance"] += sum(t[k] for k in ("adv_entry", "adv_copy_dot", "adv_dcsrch", "adv_projgr_tests", "adv_ls_step", "adv_build_B", "adv_cauchy", "adv_subsm", "adv_ls_start"))
        per_closure["advance"] = round(t["advance"] / F, 1)
        tot = sum(t[k] for k in ("load", "gram", "sweep_A", "sweep_B", "sweep_C", "gradient", "reduce", "advance",
                                 "factorise", "pool_pass"))
        out[f"n={n},B={B}"] = {"B": B, "launch_ms": e0.elapsed_time(e1), "closures_per_fit": F / C,
                         "sweep_blocks_per_closure": t["sweep_blocks"] / F,
                         "cycles_per_closure": per_closure, "advance_split": adv, "cycles_per_campaign": per_campaign,
                         "share": {k: round(t[k] / tot, 4) for k in ("gram", "sweep_A", "sweep_B", "sweep_C", "gradient",
                                                                       "reduce", "advance", "factorise", "pool_pass", "load")}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
