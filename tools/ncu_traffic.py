#!/usr/bin/env python
"""profiles/traffic.json: DRAM bytes per launch of every kernel captured with `ncu --set full` (bench.py's roofline.traffic).
usage: ncu_traffic.py name=report.ncu-rep [name=report.ncu-rep ...]   (name = the bench's kernel label)"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.path.join(ROOT, "profiles", "traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
for arg in sys.argv[1:]:
    name, rep = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, d = rows[0], rows[1], rows[2]
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = h.index(k)
        tot += float(d[i].replace(",", "")) * scale.get(u[i], 1.0)
    out[name] = {"bytes": tot, "source": os.path.relpath(rep, ROOT), "kernel": d[h.index("Kernel Name")]}
    print(name, tot / 1e6, "MB")
json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)
