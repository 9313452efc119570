"""Summarise an .ncu-rep (read here, no GPU needed) into a small CSV for profiles/: one row per profiled launch with
the metrics the roofline discussion uses.  Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.csv"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [w for w in WANT if w in idx]
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(cols)
        w.writerow([units[idx[c]] for c in cols])
        for r in rows[2:]:
            w.writerow([r[idx[c]] for c in cols])
    print("wrote", out, len(rows) - 2, "launches")


if __name__ == "__main__":
    main()
