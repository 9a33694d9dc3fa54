"""Condense an .ncu-rep (one kernel launch, `ncu --set full`) into the text summary committed under profiles/:
the launch/throughput metrics the bench's roofline block quotes and the stall-reason shares of the PC samples.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/rNN_<name>.txt"""
import csv, io, subprocess, sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit", "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "sm__inst_executed.avg.per_cycle_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__sass_inst_executed_op_shared", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__inst_executed_pipe_tensor")

def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, vals = rows[0], rows[1], rows[2]
    print(f"# {path}: kernel {vals[head.index('Kernel Name')][:80]}  grid {vals[head.index('Grid Size')]}  block {vals[head.index('Block Size')]}")
    stalls = {}
    for h, u, v in zip(head, units, vals):
        if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued"):
            try: stalls[h] = float(v.replace(",", ""))
            except ValueError: pass
        elif any(h.startswith(k) for k in KEEP):
            print(f"{h} [{u}] = {v}")
    tot = sum(stalls.values()) or 1.0
    for h, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]:
        print(f"{100 * v / tot:6.2f}%  {h}")

if __name__ == "__main__":
    main(sys.argv[1])
