#!/usr/bin/env python
"""Per-source-line summary of an .ncu-rep captured with --import-source on (compile with -lineinfo).
usage: ncu_lines.py report.ncu-rep [top_n]   -> lines ranked by stall samples, with executed warp instructions"""
import csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, hdr = None, None
items = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr and cur and r[0].isdigit():
        si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall = {hdr[k]: int(r[k]) for k in range(len(hdr)) if hdr[k].startswith("stall_") and "Not Issued" not in hdr[k] and k < len(r) and r[k].isdigit() and int(r[k])}
        try:
            items.append((int(r[si] or 0), int(r[ii] or 0), cur, r[0], r[1].strip()[:100], stall))
        except ValueError:
            pass
tot_s = sum(i[0] for i in items); tot_i = sum(i[1] for i in items)
print(f"total samples {tot_s}  total warp instructions {tot_i}")
by_ins = "--by-instructions" in sys.argv
agg = {}
for it in items:
    for k, v in it[5].items():
        agg[k] = agg.get(k, 0) + v
print("stalls:", ", ".join(f"{k[6:]} {v/max(tot_s,1):.1%}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for s, i, f, ln, src, stall in sorted(items, key=(lambda x: (x[1], x[0])) if by_ins else None, reverse=True)[:top]:
    st = ",".join(f"{k[6:]}:{v}" for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:3])
    print(f"{s/max(tot_s,1):6.1%} smp {i/max(tot_i,1):6.1%} ins  {f}:{ln:>4s}  {src}   [{st}]")
