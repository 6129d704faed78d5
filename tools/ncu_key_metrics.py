"""Print the handful of ncu raw-page metrics used in profiles/*.md for every kernel of a report.
usage: python tools/ncu_key_metrics.py report.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_stores", "sass__inst_executed_local_loads"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("==", d.get("Kernel Name", "?")[:70])
        for k in KEYS:
            if k in d:
                print("  %-68s %s %s" % (k, d[k], u.get(k, "")))
        st = sorted(((float(v), k[len(STALL):-len("_per_issue_active.ratio")]) for k, v in d.items()
                     if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v), reverse=True)
        print("  stalls per issue:", ", ".join("%s %.2f" % (k, v) for v, k in st[:7]))


if __name__ == "__main__":
    main()
