#!/usr/bin/env python
"""tools/ncu_summary.py REPORT.ncu-rep [...]: the handful of counters DESIGN.md and profiles/*.md quote,
from `ncu -i REPORT --page raw --csv` (run where ncu is installed; no GPU needed)."""
import csv, io, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        print(f"## {rep}: {name}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"| {w} | {vals[i]} {units[i]} |")
