#!/usr/bin/env python
"""Eager (no CUDA graph) plans of the bench workload for profiling under ncu:
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/prof_plan.py
    ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 -o gpurun_out/rollout python tools/prof_plan.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from mppi_isaac_b200 import MPPIisaacPlanner  # noqa: E402
from mppi_isaac_b200.objectives import PandaReachObjective  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else bench.K_PER_GPU
plans = int(sys.argv[2]) if len(sys.argv) > 2 else 4
planner = MPPIisaacPlanner(bench.panda_cfg(K, "cuda:0"), PandaReachObjective(), use_cuda_graph=False)
q0, goal = bench.synthetic_state()
planner.sim.set_actor_position_by_name(goal, "goal")
planner.sim.reset_robot_state(q0, np.zeros(7))
for _ in range(plans):
    planner.mppi.command()
torch.cuda.synchronize()
print("done", planner.mppi._action.cpu().numpy())
