"""Re-runs the variational oracle against every VarGP SMOKE trace of the reference (15 campaigns x 40 rows) and
the teacher-forced VarTGP / VarLGP checks of SURVEY App. C.5, and writes
tests/golden/oracle_kat_report_Var_smoke.json (committed; tests/test_oracle_var_golden.py re-runs the pinned
subset).  ~2 min on 8 cores:  python tools/validate_oracle_var_smoke.py"""
import json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _varpin
from _fixtures import DATASETS, GOLDEN, trace

if __name__ == "__main__":
    t0 = time.time()
    jobs = [(name, seed) for name in DATASETS for seed in (45, 359, 6185)]
    mc = [(k, n, s, 0, 0) for k in ("VarTGP", "VarLGP") for n, s in (("AutoAM", 45), ("AgNP", 45), ("Perovskite", 359))]
    v, m = _varpin.run_all(jobs, mc)
    rep = {"settings": _varpin.SMOKE, "VarGP_first_divergence": {}, "VarGP_exact": [], "teacher_forced": {}}
    for name, seed, sel in v:
        g = trace("smoke", "VarGP", name, seed)
        fd = next((i for i, (a, b) in enumerate(zip(sel, g)) if a != b), len(g))
        rep["VarGP_first_divergence"][f"{name}/{seed}"] = fd
        if sel == g:
            rep["VarGP_exact"].append(f"{name}/{seed}")
    for kind, name, seed, ranks, gaps in m:
        rep["teacher_forced"][f"{kind}/{name}/{seed}"] = {"rank_first": sum(r == 0 for r in ranks), "steps": len(ranks),
                                                          "worst_rank": max(ranks), "max_gap": max(gaps), "ranks": ranks}
    rep["wall_s"] = round(time.time() - t0, 1)
    json.dump(rep, open(os.path.join(GOLDEN, "oracle_kat_report_Var_smoke.json"), "w"), indent=1)
    print(json.dumps(rep, indent=1))
