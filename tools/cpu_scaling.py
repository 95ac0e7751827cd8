#!/usr/bin/env python
"""Thread scaling of the CPU restatement's rollout (oracle_rollout, panda reach K = 4096, T = 30) on this host:
    python tools/cpu_scaling.py
prints rollout-steps/s per thread count -- how many cores the box really gives this process."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from oracle import oracle as orc  # noqa: E402
from scenes import panda_setup  # noqa: E402

K, T = 4096, 30
sc, p, s0 = panda_setup(K=K, T=T)
a = np.random.default_rng(0).uniform(-0.2, 0.2, (T, 7, K)).astype(np.float32)
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
base = None
for nth in (1, 2, 4, 8, 16, 32, 64, 96, 128, 192):
    if nth > 2 * (os.cpu_count() or 1):
        break
    orc.rollout(sc.model, p, s0, a, nthreads=nth, want_obs=True)
    t0 = time.perf_counter()
    n = 1 if nth < 4 else 3
    for _ in range(n):
        orc.rollout(sc.model, p, s0, a, nthreads=nth, want_obs=True)
    dt = (time.perf_counter() - t0) / n
    rate = K * T / dt
    base = base or rate
    print(f"threads {nth:4d}: {dt * 1e3:8.1f} ms  {rate:.3e} rollout-steps/s  speed-up {rate / base:6.1f}  efficiency {rate / base / nth:.2f}", flush=True)
