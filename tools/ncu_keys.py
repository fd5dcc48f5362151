#!/usr/bin/env python
"""Print the handful of ncu raw-page metrics we look at (usage: ncu_keys.py report.ncu-rep [kernel-row])."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "gpc__cycles_elapsed.max.per_second", "sm__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64",
        "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "launch__registers_per_thread", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp", "smsp__warp_issue_stalled"]
for r in rows[2:]:
    print("-" * 100)
    for h, u, v in zip(hdr, units, r):
        if any(h == k or (k.endswith("_") and h.startswith(k)) or (k in ("smsp__average_warp", "smsp__warp_issue_stalled", "launch__occupancy_limit", "sm__inst_executed_pipe_fp64") and h.startswith(k)) for k in KEYS):
            if v not in ("", "0"):
                print(f"{h:95s} {u:12s} {v}")
