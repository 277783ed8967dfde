#!/usr/bin/env python
"""Prints the metrics quoted in profiles/*.md from an ncu report: python tools/ncu_summary.py report.ncu-rep
(runs `ncu -i ... --page raw --csv` and picks the columns; one block per captured launch)."""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_l1tex2xbar_write_bytes.sum", "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]
text = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(text)))
header, units = rows[0], rows[1]
stalls = [h for h in header if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h]
for row in rows[2:]:
    d = dict(zip(header, row))
    u = dict(zip(header, units))
    print("----", d["Kernel Name"][:110])
    for name in WANT:
        if name in d:
            print("   %-75s %s %s" % (name, d[name], u[name]))
    ranked = sorted(((float(d[s].replace(",", "")) if d[s] else 0.0, s) for s in stalls), reverse=True)[:6]
    for value, name in ranked:
        print("   stall %-69s %.2f" % (name.split("issue_stalled_")[1].split("_per")[0], value))
