#!/usr/bin/env python
"""Per-source-line totals from an ncu report:  python tools/ncu_lines.py report.ncu-rep [top]
(reads `ncu --page source --print-source cuda,sass --csv`; prints samples and executed warp instructions per CUDA line)."""
import csv
import subprocess
import sys
from collections import defaultdict

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fpath, hdr, lines = None, None, {}
for r in rows:
    if r and r[0] == "File Path":
        fpath = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
    elif hdr and len(r) == len(hdr) and r[0]:
        try:
            lines[(fpath, int(r[0]))] = (r[1], int(r[i_s] or 0), int(r[i_i] or 0))
        except ValueError:
            pass
ts, ti = sum(v[1] for v in lines.values()), sum(v[2] for v in lines.values())
print(f"total samples {ts}  total warp instructions {ti}")
perfile = defaultdict(lambda: [0, 0])
for (f, _), v in lines.items():
    perfile[f][0] += v[1]; perfile[f][1] += v[2]
for f, v in perfile.items():
    print(f"  {f}: samples {100*v[0]/max(ts,1):.1f}%  instr {100*v[1]/max(ti,1):.1f}%")
for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*v[1]/max(ts,1):5.1f}% smp {100*v[2]/max(ti,1):5.1f}% ins  {f}:{ln}  {v[0].strip()[:110]}")
