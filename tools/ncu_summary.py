#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel, --set full) into a small text file for profiles/."""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "sm__cycles_elapsed.avg", "smsp__cycles_elapsed.avg.per_second"]

def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# summary of {rep} (ncu --set full --clock-control none), one block per captured launch"]
    for d in data:
        lines.append(f"kernel: {d[idx['Kernel Name']]}")
        for k in KEYS:
            if k in idx:
                lines.append(f"  {k} [{units[idx[k]]}] = {d[idx[k]]}")
        stalls = []
        for h, i in idx.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(d[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        lines.append("  stalls per issued instruction (top 8): " + ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]))
        try:
            rd, wr = float(d[idx["dram__bytes_read.sum"]]), float(d[idx["dram__bytes_write.sum"]])
            lines.append(f"  dram traffic per launch = {rd + wr:.1f} {units[idx['dram__bytes_read.sum']]}")
        except Exception:
            pass
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
