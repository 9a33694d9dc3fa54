"""Per-kernel SASS opcode summary of the shipped library (cuobjdump -sass): instruction count and how many of them
are fp64 tensor-core MMAs (DMMA), fp64 FMA / ADD / MUL, warp shuffles, barriers, shared / global / local memory
accesses and asynchronous copies.  Evidence for DESIGN.md section 5 ("which pipe does the work").
usage: python tools/sass_opcount.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "stpbenchmark_b200", "libstp_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = re.findall(r"arch = (sm_\w+)", out)
CLASSES = [("DMMA", r"^DMMA"), ("DFMA", r"^DFMA"), ("DADD", r"^DADD"), ("DMUL", r"^DMUL"), ("MUFU", r"^MUFU"), ("SHFL", r"^SHFL"),
           ("BAR", r"^BAR"), ("LDS", r"^LDS"), ("STS", r"^STS"), ("LDG", r"^LDG(?!STS)"), ("STG", r"^STG"), ("LDGSTS", r"^LDGSTS"),
           ("LDL/STL", r"^(LDL|STL)"), ("ATOM", r"^(ATOM|RED)"), ("UTC*MMA/UTMA", r"^(UTC|UTMA|LDTM|STTM)")]
cur, counts, total = None, {}, collections.Counter()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = m.group(1)
        k = re.search(r"\d+(k_[a-z0-9_]+?)(?:I|E)", name)
        cur = (k.group(1) if k else name) + ("<BIG>" if "ILb1E" in name else ("<small>" if "ILb0E" in name else ""))
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for cname, pat in CLASSES:
            if re.match(pat, op):
                counts[cur][cname] += 1
print(f"# {os.path.relpath(lib, ROOT)}: architectures {sorted(set(arch))}")
print(f"{'kernel':24s} {'instr':>7s} " + " ".join(f"{c:>7s}" for c, _ in CLASSES))
for k in sorted(counts, key=lambda k: -total[k]):
    print(f"{k:24s} {total[k]:7d} " + " ".join(f"{counts[k][c]:7d}" for c, _ in CLASSES))
print("# fp64 has no tcgen05 / UMMA form: the tensor-core instruction of this library is DMMA (mma.sync.m8n8k4.f64);")
print("# LDGSTS = cp.async (operand staging of the n_cap > 144 variational contractions).")
