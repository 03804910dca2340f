"""Condenses `ncu -i <rep> --page raw --csv` into the per-launch summary kept under profiles/.
usage: ncu -i gpurun_out/prof.ncu-rep --page raw --csv | python tools/ncu_summary.py > profiles/<name>.csv"""
import csv, sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
]
rows = list(csv.reader(sys.stdin))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units, launches = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
col = {n: i for i, n in enumerate(names)}
out = csv.writer(sys.stdout)
out.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(launches))])
out.writerow(["Kernel Name", ""] + [r[col["Kernel Name"]] for r in launches])
stall = [n for n in names if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("_per_issue_active.ratio")]
for n in KEEP + stall:
    if n in col:
        out.writerow([n, units[col[n]]] + [r[col[n]] for r in launches])
