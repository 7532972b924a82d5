"""Print the headline metrics of the first kernel in an .ncu-rep (raw page): duration, DRAM bytes/throughput, L2 hit rate,
occupancy, issue utilisation, stall breakdown.  usage: python tools/ncu_summary.py report.ncu-rep [kernel-index]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2 + idx]
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active"]
for h, u, v in zip(hdr, units, vals):
    if h in KEYS or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") \
            or "pipe_tensor" in h and "pct_of_peak_sustained_active" in h and h.endswith(".avg.pct_of_peak_sustained_active"):
        try:
            if float(v.replace(",", "")) == 0 and "stalled" in h:
                continue
        except ValueError:
            pass
        print(f"{h:100s} {u:12s} {v}")
