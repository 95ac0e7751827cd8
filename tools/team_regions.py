#!/usr/bin/env python
"""Share of the team kernel's warp instructions / samples per code region, from an ncu report with source (--import-source on):
    python tools/team_regions.py report.ncu-rep"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "mppi_isaac_b200", "csrc", "rollout_team.cu")).read().splitlines()


def find(pat):
    for n, l in enumerate(src, 1):
        if pat in l:
            return n
    raise KeyError(pat)


marks = [(find("auto refresh_free"), "helpers (ld/st, team_sum, philox), kernel prologue"), (find("auto shapes_world"), "init / refresh_free"), (find("auto append"), "shapes_world"),
         (find("auto detect ="), "detect geometry (points / sphere / near / pairs)"), (find("auto chain_sign"), "detect loop"), (find("auto solve_contacts"), "chain_sign"),
         (find("// per contact, once"), "solve: coordinate velocities"), (find("// the sweeps:"), "solve: rows + effective masses"),
         (find("// ---- back to the bodies"), "solve: GS visits"), (find("auto integrate_free"), "solve: write-back / net force"),
         (find("auto write_obs"), "integrate_free"), (find("auto articulation ="), "write_obs"), (find("constexpr int VB"), "articulation: per-body terms, composites"),
         (find("Kin kn;"), "articulation: joint-space LDL"), (10 ** 6, "main loop: targets, hand-over, integrate, stores")]
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), sys.argv[1], "2000"], capture_output=True, text=True).stdout
tot, smp = collections.Counter(), collections.Counter()
for l in out.splitlines():
    mt = re.match(r"\s*([\d.]+)% smp\s+([\d.]+)% ins\s+(\S+):(\d+)", l)
    if not mt:
        continue
    f, ln = mt.group(3), int(mt.group(4))
    key = "inlined vector math / shuffles (headers)" if f != "rollout_team.cu" else next(name for mk, name in marks if ln < mk)
    tot[key] += float(mt.group(2)); smp[key] += float(mt.group(1))
print(out.splitlines()[0])
for k, v in tot.most_common():
    print(f"{v:6.1f}% ins {smp[k]:6.1f}% smp  {k}")
