#!/usr/bin/env python
"""One line per launch of an .ncu-rep (`--set full`): duration, DRAM bytes, launch shape, instructions, issue/occupancy.

    python profiles/ncu_table.py gpurun_out/step_final.ncu-rep > profiles/r1_ncu_step_final.txt
"""
import csv
import subprocess
import sys

COLS = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "warp_inst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"), ("lts__t_sector_hit_rate.pct", "l2_hit_%")]

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = [hdr.index(c) for c, _ in COLS]
print(" | ".join(f"{n} [{units[i]}]" if units[i] else n for (_, n), i in zip(COLS, idx)))
for r in rows[2:]:
    print(" | ".join(r[i].split("(")[0][:36] if k == 0 else r[i] for k, i in enumerate(idx)))
