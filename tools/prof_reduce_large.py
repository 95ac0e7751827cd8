#!/usr/bin/env python
"""K3 alone at a bandwidth-relevant size for ncu:  ncu --set full -k regex:mppib_reduce -s 3 -c 1 python tools/prof_reduce_large.py [K]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from mppi_isaac_b200 import MPPIisaacPlanner  # noqa: E402
from mppi_isaac_b200.backend import CudaBackend  # noqa: E402
from mppi_isaac_b200.model.blob import MppibParams  # noqa: E402
from mppi_isaac_b200.objectives import PandaReachObjective  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
planner = MPPIisaacPlanner(bench.panda_cfg(10000, "cuda:0"), PandaReachObjective(), use_cuda_graph=False)
p = MppibParams.from_buffer_copy(bytes(planner.mppi.backend.params))
p.K = K
be = CudaBackend("cuda:0")
be.create(planner.sim.scene.model, p)
T, nu = planner.mppi.T, planner.mppi.nu
bufs = [(torch.randn((T, nu, K), device="cuda") * 0.3, torch.rand((T, K), device="cuda") * 10) for _ in range(3)]   # 3 x 252 MB > L2
U, partial = torch.zeros((T, nu), device="cuda"), torch.zeros(2 + T * nu, device="cuda")
for i in range(6):
    x, c = bufs[i % 3]
    be.reduce(c, x, U, partial)
torch.cuda.synchronize()
print("done", 4 * K * T * (nu + 1) + 4 * (T * nu + 2), "algorithmic bytes per launch")
