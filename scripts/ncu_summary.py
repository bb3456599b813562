"""Reduce an ncu --set full report to the metrics DESIGN.md / bench.py cite.

    python scripts/ncu_summary.py gpurun_out/r01_full.ncu-rep profiles/r01_ncu_full_summary.csv

Runs here (no GPU needed): `ncu -i <rep> --page raw --csv` and keeps one row per profiled launch.
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "lts__t_sectors_op_write.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    print(f"{len(rows) - 2} launches -> {out}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
