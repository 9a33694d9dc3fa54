"""Static code footprint of one kernel by source file / region, from `nvdisasm -gi` of the built library:
how many SASS instructions (16 B each) every header contributes.  A single warp that runs straight through tens of
KB of cold code is instruction-fetch bound (L0 6 KB, L1.5 32 KB per SM): this is the tool that found the L-BFGS-B
advance of k_exact_bo_step at 370 KB of inlined code.
usage: python tools/sass_footprint.py [kernel-substring] [lib.so]"""
import collections, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    kern = sys.argv[1] if len(sys.argv) > 1 else "k_exact_bo_stepE"
    lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "stpbenchmark_b200", "libstp_b200.so")
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, capture_output=True)
        cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
        text = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    infun, first, by_file, by_line = False, None, collections.Counter(), collections.Counter()
    for ln in text.splitlines():
        if ln.startswith("\t.section\t.text."):
            infun = kern in ln
            continue
        if ln.startswith("\t.section"):
            infun = False
            continue
        if not infun:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            if first is None:                       # innermost frame comes first, then its "inlined at" chain
                first = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            if first:
                last = first
            by_file[last[0]] += 1
            by_line[last] += 1
            first = None
    tot = sum(by_file.values())
    print(f"{kern}: {tot} SASS instructions = {tot * 16 / 1024:.1f} KB")
    for f, c in by_file.most_common():
        print(f"  {f:32s} {c:7d}  {c * 16 / 1024:7.1f} KB")
    regions = {"lbfgsb.cuh": [(107, 177, "build_B"), (178, 281, "cauchy"), (282, 395, "subsm"), (396, 459, "dcstep"),
                              (460, 507, "dcsrch"), (508, 536, "init"), (537, 622, "begin_iteration"),
                              (623, 760, "advance_once")]}
    for f, regs in regions.items():
        acc = collections.Counter()
        for (ff, l), c in by_line.items():
            if ff == f:
                acc[next((n for lo, hi, n in regs if lo <= l <= hi), "other")] += c
        if acc:
            print(f"  regions of {f}: " + ", ".join(f"{n} {c}" for n, c in acc.most_common()))


if __name__ == "__main__":
    main()
