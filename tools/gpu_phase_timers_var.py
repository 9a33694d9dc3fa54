"""Per-phase cycle budget of var_epoch inside k_var_fit (thread 0's clock64 between the phase marks).
Builds a -DSTP_PHASE_TIMERS variant of the library beside the shipped one and runs E epochs of B VarTGP campaigns of
n points.  usage (under gpurun): python tools/gpu_phase_timers_var.py [cfg5] > gpurun_out/var_phase_timers.json"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANT = os.path.join(ROOT, "stpbenchmark_b200", "libstp_b200_timers.so")
from stpbenchmark_b200 import build as b
if not os.environ.get("STP_LIB"):     # STP_LIB: a prebuilt -DSTP_PHASE_TIMERS variant (A/B runs)
    subprocess.run([b.nvcc_path()] + b.NVCC_FLAGS + ["-DSTP_PHASE_TIMERS", "-o", VARIANT] + b.SOURCES, check=True)
    os.environ["STP_LIB"] = VARIANT
import torch
from stpbenchmark_b200 import _lib, synthetic, varops
CFG5 = "cfg5" in sys.argv
dev = "cuda"
X, y = synthetic.cfg5_pool(0) if CFG5 else synthetic.cfg4_pool(0)
d = X.shape[1]; N = X.shape[0]
out = {}
for B, n, E in ([(148, 256, 10), (296, 256, 10)] if CFG5 else [(296, 30, 40), (296, 60, 40), (296, 89, 40), (1024, 50, 40)]):
    cap = n + 1
    g = torch.Generator().manual_seed(n)
    Xb = torch.zeros(B, cap, d, dtype=torch.double); yb = torch.zeros(B, cap, dtype=torch.double)
    for bb in range(B):
        idx = torch.randperm(N, generator=g)[:n]; Xb[bb, :n] = X[idx]; yb[bb, :n] = y[idx]
    Xb, yb = Xb.to(dev), yb.to(dev); nb = torch.full((B,), n, dtype=torch.int32, device=dev)
    st = varops.alloc_state(B, cap, d, dev); status = torch.zeros(B, dtype=torch.int32, device=dev); work = _lib.var_workspace(B, cap, dev)
    args = (Xb, yb, nb, st, 1, E, 0.01, 0.01, varops.var_hyp0(d), varops.var_prior_vector(), varops.gauss_hermite_table(), status, work)
    _lib.var_fit(*args, fresh=True, seed=1); torch.cuda.synchronize()
    _lib.phase_timers(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _lib.var_fit(*args, fresh=True, seed=1); e1.record(); torch.cuda.synchronize()
    t = _lib.phase_timers()
    ep = max(t["epochs"], 1)
    names = [v for k, v in sorted(_lib.VAR_PHASE_NAMES.items()) if v != "epochs"] + ["sweep_A", "sweep_B", "sweep_C"]   # the blocked sweeps (P, K_ZZ) charge the exact path's ids
    tot = sum(t[k] for k in names) or 1
    out[f"n={n},B={B}"] = {"launch_ms": e0.elapsed_time(e1), "ms_per_epoch": e0.elapsed_time(e1) / E, "epochs_counted": ep, "bad": int((status != 0).sum()),
                           "cycles_per_epoch": {k: round(t[k] / ep) for k in names}, "share": {k: round(t[k] / tot, 4) for k in names},
                           "cycles_per_epoch_total": round(tot / ep)}
print(json.dumps(out, indent=1))
