#!/usr/bin/env python
"""K3 launch-shape sweep (cold L2, graph-timed): CTA count and tile width at the BASELINE K.   python tools/tune_reduce.py [K ...]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Ks = [int(a) for a in sys.argv[1:]] or [10000, 65536]
code = ("import sys; sys.path.insert(0, %r); import bench, torch, numpy as np\n"
        "from mppi_isaac_b200 import MPPIisaacPlanner\nfrom mppi_isaac_b200.objectives import PandaReachObjective\n"
        "p = MPPIisaacPlanner(bench.panda_cfg(10000, 'cuda:0'), PandaReachObjective(), use_cuda_graph=False)\n"
        "r = bench.k3_roofline(p, 6486.5, 'x', %r)\n"
        "print('RES', ' '.join('%%d:%%.2f' %% (e['K'], e['us']) for e in r))\n") % (ROOT, Ks)
res = {}
for grid in ("", "37", "74", "111"):
    for wide in ("", "1"):
        env = dict(os.environ)
        if grid:
            env["MPPIB_K3_GRID"] = grid
        if wide:
            env["MPPIB_K3_WIDE"] = wide
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("RES")]
        res[f"grid={grid or 'sms'} wide={wide or 'auto'}"] = line[0][4:] if line else out.stderr[-200:]
        print(f"grid={grid or 'sms':4s} wide={wide or 'auto':4s}  {res[f'grid={grid or chr(115)+chr(109)+chr(115)} wide={wide or chr(97)+chr(117)+chr(116)+chr(111)}']}", flush=True)
print(json.dumps(res))
