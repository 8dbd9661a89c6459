#!/usr/bin/env python
"""Text summary of one `ncu --set full` report for profiles/: duration, DRAM traffic, pipe utilisation, occupancy, stalls, hottest lines.
usage: ncu_summary.py report.ncu-rep [algorithmic_bytes]"""
import csv, io, subprocess, sys, os

rep = sys.argv[1]
alg = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
for d in rows[2:]:
    g = lambda k: d[h.index(k)] if k in h else "n/a"
    f = lambda k: float(g(k).replace(",", "")) if g(k) not in ("n/a", "") else float("nan")
    unit = lambda k: u[h.index(k)] if k in h else ""
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
    dr = f("dram__bytes_read.sum") * scale.get(unit("dram__bytes_read.sum"), 1.0)
    dw = f("dram__bytes_write.sum") * scale.get(unit("dram__bytes_write.sum"), 1.0)
    t = f("gpu__time_duration.sum") * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(unit("gpu__time_duration.sum"), 1e-6)
    print(f"kernel            {g('Kernel Name')}   grid {g('launch__grid_size')} x {g('launch__block_size')}  regs {g('launch__registers_per_thread')}")
    print(f"duration          {t*1e6:.2f} us (under ncu: cold caches, serialised)")
    print(f"dram traffic      read {dr/1e6:.2f} MB + write {dw/1e6:.2f} MB = {(dr+dw)/1e6:.2f} MB  -> {(dr+dw)/t/1e9:.0f} GB/s" +
          (f"   algorithmic {alg/1e6:.2f} MB (traffic/alg = {(dr+dw)/alg:.2f})" if alg else ""))
    for k in ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
              "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "launch__waves_per_multiprocessor",
              "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
              "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"):
        print(f"{k:62s} {g(k):>16s} {unit(k)}")
tool = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_lines.py")
print(subprocess.run([sys.executable, tool, rep, "14"], capture_output=True, text=True).stdout)
