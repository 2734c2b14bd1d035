#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/:
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_xxx.txt "free-form note"
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["# " + rep, "# " + note, "# ncu --set full --clock-control none (one launch per kernel; cold-ish caches)"]
    for r in rows[2:]:
        lines.append("")
        lines.append("kernel: " + r[hdr.index("Kernel Name")][:150])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append("  %-68s %18s %s" % (k, r[i], units[i]))
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        lines.append("  top stalls (warps per issue): " + ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)[:6]))
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
