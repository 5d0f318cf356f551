#!/usr/bin/env python3
"""Key metrics per captured launch of an .ncu-rep (or of its `--page raw --csv` export): ncu_raw.py report.ncu-rep|raw.csv"""
import csv, subprocess, sys, io
out = (open(sys.argv[1]).read() if sys.argv[1].endswith(".csv") else
       subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed.sum", "inst"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wf"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conf"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("sm__cycles_elapsed.max", "cyc"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed","smem%"),
        ("smsp__inst_executed.sum","inst2")]
for r in rows[2:]:
    print("== " + r[idx["Kernel Name"]][:80])
    print("   " + ", ".join(f"{n}={r[idx[m]]}{units[idx[m]] if n in ('time','rd','wr') else ''}" for m, n in want if m in idx))
