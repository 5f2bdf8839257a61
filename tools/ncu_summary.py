"""Condense an `ncu --set full` report into the per-launch columns profiles/ keeps.
usage: ncu -i report.ncu-rep --page raw --csv > raw.csv; python tools/ncu_summary.py raw.csv > summary.csv"""
import csv, sys
COLS = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max"]
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
idx = [hdr.index(c) for c in COLS if c in hdr]
w = csv.writer(sys.stdout)
for r in rows:
    w.writerow([r[i] for i in idx])
