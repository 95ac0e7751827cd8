#!/usr/bin/env python
"""K3 A/B on the GPU box: the warp-specialised reduce kernel (default) against the block-synchronous one (MPPIB_K3_VARIANT),
cold L2 (inputs rotated over > 2x the L2), graph-timed; also checks that both produce the same shard row.

    python tools/k3_ab.py [K ...]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Ks = [int(a) for a in sys.argv[1:]] or [10000, 65536, 131072, 262144]
code = ("import sys, json; sys.path.insert(0, %r); import bench, torch, numpy as np\n"
        "from mppi_isaac_b200 import MPPIisaacPlanner\nfrom mppi_isaac_b200.objectives import PandaReachObjective\n"
        "from mppi_isaac_b200.backend import CudaBackend\nfrom mppi_isaac_b200.model.blob import MppibParams\n"
        "p = MPPIisaacPlanner(bench.panda_cfg(10000, 'cuda:0'), PandaReachObjective(), use_cuda_graph=False)\n"
        "peak, _ = bench.measured_peak_gbs()\n"
        "r = bench.k3_roofline(p, peak, %r)\n"
        "pp = MppibParams.from_buffer_copy(bytes(p.mppi.backend.params)); pp.K = 65536\n"
        "be = CudaBackend('cuda:0'); be.create(p.sim.scene.model, pp)\n"
        "g = torch.Generator(device='cuda').manual_seed(0)\n"
        "x = torch.randn((30, 7, 65536), device='cuda', generator=g) * 0.3; c = torch.rand((30, 65536), device='cuda', generator=g) * 10\n"
        "U = torch.zeros((30, 7), device='cuda'); part = torch.zeros(212, device='cuda')\n"
        "be.reduce(c, x, U, part); torch.cuda.synchronize()\n"
        "print('RES', json.dumps({'sweep': r, 'row': part.cpu().tolist()}))\n") % (ROOT, Ks)
out = {}
for name, env in (("warp-specialised", {}), ("block-synchronous", {"MPPIB_K3_VARIANT": "auto"})):
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("RES")]
    if not line:
        print(name, "FAILED", r.stderr[-800:])
        continue
    out[name] = json.loads(line[0][4:])
print("| K | bytes (MB) | " + " | ".join(f"{n} us / GB/s / frac" for n in out) + " |")
print("|---|---|" + "---|" * len(out))
for i, K in enumerate(Ks):
    cells = []
    for n in out:
        e = out[n]["sweep"][i]
        cells.append(f"{e['us']:.2f} / {e['GBps']:.0f} / {e['frac']:.3f}")
    print(f"| {K} | {out[list(out)[0]]['sweep'][i]['bytes'] / 1e6:.1f} | " + " | ".join(cells) + " |")
if len(out) == 2:
    a, b = (out[n]["row"] for n in out)
    d = max(abs(u - v) / max(1.0, abs(v)) for u, v in zip(a, b))
    print(f"\nshard row (beta, eta, W) at K = 65536: max relative difference between the two kernels = {d:.2e}")
print(json.dumps({n: out[n]["sweep"] for n in out}))
