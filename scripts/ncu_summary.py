"""Summarise an ncu report (.ncu-rep, read here on the CPU box with `ncu -i`) into the per-kernel CSV kept under profiles/.
  python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/<name>.csv"""
import csv, io, subprocess, sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
keep = [m for m in METRICS if m in idx]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["Kernel Name"] + keep)
    w.writerow([""] + [units[idx[m]] for m in keep])
    for r in data:
        w.writerow([r[idx["Kernel Name"]]] + [r[idx[m]] for m in keep])
print(open(out).read())
