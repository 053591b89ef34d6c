"""Prints the metrics we track from an .ncu-rep (run locally: ncu can read reports without a GPU)."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
KEEP = ["gpu__time_duration.sum", "launch__registers_per_thread ", "launch__grid_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum ", "dram__bytes_write.sum ", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum ", "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled",
        "sass__inst_executed_local", "sm__cycles_elapsed.avg ", "smsp__inst_executed_op_", "sm__sass_inst_executed_op_"]
for h, u, v in zip(hdr, units, vals):
    if any((h + " ").startswith(k) for k in KEEP):
        if "issue_stalled" in h and float(v or 0) < 0.05:
            continue
        print(f"{h} [{u}] = {v}")
