"""Summarise an .ncu-rep (ncu --set full capture) into a small committed text file under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_attn.ncu-rep profiles/r01_attn_v0_msn_b16.txt "note"
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = ["# ncu --set full --clock-control none summary of %s" % rep, "# " + note, ""]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        lines.append("kernel: %s  grid=%s block=%s" % (d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
        for k in KEYS:
            if k in d:
                lines.append("  %-82s %14s %s" % (k, d[k], u[k]))
        st = sorted(((float(d[k].replace(",", "")), k) for k in hdr if k.startswith(STALL) and k.endswith("_per_issue_active.ratio")),
                    reverse=True)
        lines.append("  top stall reasons (warps stalled per issue-active cycle):")
        for v, k in st[:7]:
            lines.append("    %-40s %8.3f" % (k[len(STALL):-len("_per_issue_active.ratio")], v))
        if "dram__bytes_read.sum" in d:
            lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
