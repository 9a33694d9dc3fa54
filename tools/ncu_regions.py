"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by (file, source line) and by coarse
line ranges: which region of which header the stall samples / executed instructions of a kernel fall in.
The export attributes a SASS instruction to EVERY source line of its inline stack (a shuffle of sum_vec shows up
under team.cuh and under three lines of sm_30_intrinsics.hpp), so the shares are INCLUSIVE of inlined callees and can
add up to more than 100 %.  Pass the kernel's true totals (sum over the rows of the `--print-source sass` export) as
`--totals SAMPLES INSTRUCTIONS`; without them the sum over source lines is used (over-counts deep inline stacks).
usage: python tools/ncu_regions.py export.csv [top_n] [--totals SAMPLES INSTRUCTIONS]"""
import csv, sys, collections

def load(path):
    cur = None; header = None
    per_line = collections.defaultdict(lambda: [0, 0])   # (file, line) -> [samples, instructions]
    src = {}
    for row in csv.reader(open(path, newline='')):
        if not row: continue
        if row[0] == 'File Path': cur = row[1]; header = None; continue
        if row[0] == 'Function Name': continue
        if row[0] == 'Line No': header = row; continue
        if header is None or cur is None: continue
        try: line = int(row[0])
        except ValueError: continue
        rec = dict(zip(header[2:], row[2:]))   # columns after (Line No, Source) belong to the SASS rows
        src[(cur, line)] = row[1]
        if not rec.get('Address'): continue
        try:
            s = int(rec.get('# Samples') or 0); ins = int(rec.get('Instructions Executed') or 0)
        except ValueError: continue
        per_line[(cur, line)][0] += s; per_line[(cur, line)][1] += ins
    return per_line, src

if __name__ == '__main__':
    argv = list(sys.argv)
    totals = None
    if '--totals' in argv:
        k = argv.index('--totals'); totals = (int(argv[k + 1]), int(argv[k + 2])); del argv[k:k + 3]
    per_line, src = load(argv[1]); top = int(argv[2]) if len(argv) > 2 else 40
    ts = sum(v[0] for v in per_line.values()); ti = sum(v[1] for v in per_line.values())
    if totals:
        ts, ti = totals
    print(f'total samples {ts}  warp instructions {ti}   (shares are inclusive of inlined callees)')
    by_file = collections.defaultdict(lambda: [0, 0])
    for (f, l), v in per_line.items(): by_file[f][0] += v[0]; by_file[f][1] += v[1]
    for f, v in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
        print(f'{100*v[0]/ts:6.2f}% samples {100*v[1]/max(ti,1):6.2f}% instr  {f}')
    print()
    for (f, l), v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f'{100*v[0]/ts:6.2f}% {100*v[1]/max(ti,1):6.2f}%  {f.split("/")[-1]}:{l}  {src.get((f,l),"").strip()[:110]}')
