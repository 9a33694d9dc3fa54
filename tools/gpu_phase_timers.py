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
    from stpbenchmark_b200 import _lib, synthetic
    from stpbenchmark_b200.engine import CampaignBatch
    dev = torch.device("cuda", 0)
    pools = [synthetic.cfg4_pool(g) for g in range(8)]
    PX = torch.stack([p[0] for p in pools]); PY = torch.stack([p[1] for p in pools])
    out = {}
    for n, B in [(n, B) for n in args.n for B in args.B]:
        gen = torch.Generator().manual_seed(n)
        init, pid = [], []
        for b in range(B):
            g = b % 8
            d0 = synthetic.initial_designs(pools[g][0], pools[g][1], [7000 + b], 10)[0].tolist()
            rest = [i for i in torch.randperm(2048, generator=gen).tolist() if i not in set(d0)][: n - 10]
            init.append(d0 + rest); pid.append(g)
        batch = CampaignBatch(PX, PY, pid, init, n_cap=n + 2, kind="ExactGP", device=dev)
        batch.step(); torch.cuda.synchronize()          # warm-up on a copy of the state would change n: rebuild
        batch = CampaignBatch(PX, PY, pid, init, n_cap=n + 2, kind="ExactGP", device=dev)
        _lib.phase_timers(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); batch.step(); e1.record(); torch.cuda.synchronize()
        t = _lib.phase_timers()
        F, C = max(t["closures"], 1), max(t["campaigns"], 1)
        per_closure = {k: round(t[k] / F, 1) for k in ("gram", "sweep_A", "sweep_B", "sweep_C", "gradient", "reduce", "advance")}
        per_campaign = {k: round(t[k] / C, 1) for k in ("load", "factorise", "pool_pass")}
        adv = {k: round(t[k] / F, 1) for k in ("adv_ls_step", "adv_build_B", "adv_cauchy", "adv_subsm", "adv_ls_start", "advance")}
        t["advance"] += sum(t[k] for k in ("adv_ls_step", "adv_build_B", "adv_cauchy", "adv_subsm", "adv_ls_start"))
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
