"""Summarise .ncu-rep files (gpurun_out/) into small CSVs under profiles/ (read with `ncu -i ... --page raw --csv`)."""
import csv
import io
import subprocess
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg"]


def main():
    """ncu_summary.py report.ncu-rep out.csv [kernel-name substring]: with a substring only the launches of that kernel."""
    src, dst = sys.argv[1], sys.argv[2]
    only = sys.argv[3] if len(sys.argv) > 3 else None
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    if only is not None:
        k = hdr.index("Kernel Name")
        rows = rows[:2] + [r for r in rows[2:] if only in r[k]]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none, source {src}\n")
        f.write("metric,unit," + ",".join(f"launch{i}" for i in range(len(rows) - 2)) + "\n")
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                f.write(k + "," + rows[1][i] + "," + ",".join('"%s"' % r[i] if "," in r[i] else r[i] for r in rows[2:]) + "\n")


if __name__ == "__main__":
    main()
