"""Text summary of one kernel's `ncu --set full` capture (metrics the DESIGN.md discussion uses + warp stall ratios).
usage: python tools/ncu_summary.py report.ncu-rep > profiles/xxx.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units, data = rows[0], rows[1], rows[2:]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
print(f"kernel: {data[0][h.index('Kernel Name')]}")
print(f"launches in the report: {len(data)}")
for k in keys:
    if k in h:
        i = h.index(k)
        print(f"{k:72s} {units[i]:16s} {[r[i] for r in data]}")
for i, k in enumerate(h):
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
        print(f"{k:90s} {data[-1][i]}")
