"""Summary of one kernel's `ncu --set full` capture: a text table (metrics the DESIGN.md discussion uses + warp stall
ratios) on stdout and, with --json FILE, the per-launch numbers bench.py quotes (it reads them from the committed file
and names the file in the bench line).
usage: python tools/ncu_summary.py report.ncu-rep [--json profiles/xxx.json] > profiles/xxx.txt"""
import csv, json, os, subprocess, sys
rep = sys.argv[1]
jpath = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units, data = rows[0], rows[1], rows[2:]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum",
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


def val(name, row):
    if name not in h:
        return None
    i = h.index(name)
    try:
        v = float(row[i].replace(",", ""))
    except ValueError:
        return None
    u = units[i]
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ns": 1e-3, "ms": 1e3}.get(u, 1.0)
    return v * scale


if jpath:
    n = len(data)
    avg = lambda name: (sum(val(name, r) or 0.0 for r in data) / n) if name in h else None
    js = {"source": os.path.basename(rep), "kernel": data[0][h.index("Kernel Name")], "launches": n,
          "how": "ncu --set full --clock-control none, per-launch averages (cold cache, serialised: bytes are exact, times are not bench times)",
          "dram_bytes_read": avg("dram__bytes_read.sum"), "dram_bytes_write": avg("dram__bytes_write.sum"),
          "duration_us_under_ncu": avg("gpu__time_duration.sum"),
          "tensor_pipe_pct_of_active": avg("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
          "tensor_pipe_pct_of_elapsed": avg("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
          "dram_pct_of_peak": avg("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
          "registers_per_thread": avg("launch__registers_per_thread")}
    if js["dram_bytes_read"] is not None and js["dram_bytes_write"] is not None:
        js["traffic"] = js["dram_bytes_read"] + js["dram_bytes_write"]
    json.dump(js, open(jpath, "w"), indent=1)
