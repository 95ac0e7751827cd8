#!/usr/bin/env python
"""Eager plans of a contact config for ncu:  python tools/prof_push.py [heijn|boxer|pick] [K]"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mppi_isaac_b200 import MPPIisaacPlanner, load_isaacgym_config  # noqa: E402
from mppi_isaac_b200.objectives import PandaPickObjective, PushObjective  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "heijn"
cfgname, obj, q = {
    "heijn": ("config_heijn_push_b200", PushObjective, [0.0] * 3),
    "boxer": ("config_boxer_push_b200", lambda: PushObjective(robot="boxer", link="ee_link"), [0.0, 2.5, 0.0]),
    "pick": ("config_panda_pick_b200", PandaPickObjective, [0.0, -0.94, 0.0, -2.8, 0.0, 1.8675, 0.0, 0.02, 0.02]),
}[which]
cfg = copy.deepcopy(load_isaacgym_config(cfgname))
cfg.mppi.num_samples, cfg.mppi.device = int(sys.argv[2]) if len(sys.argv) > 2 else 4000, "cuda:0"
planner = MPPIisaacPlanner(cfg, obj(), use_cuda_graph=False)
planner.sim.reset_robot_state(q, [0.0] * len(q))
for _ in range(4):
    planner.mppi.command()
torch.cuda.synchronize()
print("done", planner.mppi._action.cpu().numpy())
