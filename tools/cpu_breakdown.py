#!/usr/bin/env python
"""Where does the CPU restatement spend its time on this host? (threads sweep)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from mppi_isaac_b200 import MPPIisaacPlanner
from mppi_isaac_b200.objectives import PandaReachObjective
from oracle.backend import OracleBackend
print("cpus", os.cpu_count(), "torch threads", torch.get_num_threads())
for nth in (1, 8, 32, 64, 128):
    if nth > (os.cpu_count() or 1): break
    for tth in (1, 8):
        torch.set_num_threads(tth)
        be = OracleBackend(nthreads=nth)
        p = MPPIisaacPlanner(bench.panda_cfg(2000, "cpu"), PandaReachObjective(), backend=be)
        p.mppi.command()
        m = p.mppi
        t = {}
        def tm(name, fn, n=3):
            t0 = time.perf_counter()
            for _ in range(n): fn()
            t[name] = (time.perf_counter() - t0) / n * 1e3
        tm("sample", lambda: be.sample(0, 0, 0, 2000, m.U, None, m.actions, m.noise))
        tm("rollout", lambda: p.sim.rollout_all(m.actions))
        tm("cost", lambda: m._cost_batched())
        c = m._cost_batched()
        tm("reduce", lambda: be.reduce(c, m.noise, m.U, m.partial))
        tm("plan", lambda: m.command())
        print(f"oracle_threads={nth} torch_threads={tth}", {k: round(v, 2) for k, v in t.items()}, flush=True)
