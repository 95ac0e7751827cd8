#!/usr/bin/env python
"""Plan time of every BASELINE config (C1..C5) on ONE GPU at the per-GPU sample count of the config (C4: 16000/4, C5: 65536/8;
also the full K on one GPU where it fits).  CUDA-graph replay, CUDA events, 3 warm-up + 20 timed plans.  Prints a markdown table."""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from mppi_isaac_b200 import MPPIisaacPlanner, load_isaacgym_config  # noqa: E402
from mppi_isaac_b200.objectives import PandaPickObjective, PandaReachObjective, PointReachObjective, PushObjective  # noqa: E402

CASES = [
    ("C1 point_robot reach", "config_point_robot_b200", 128, PointReachObjective, [0.1, 0.0, 0.0]),
    ("C2 panda reach", "config_panda_b200", 1000, PandaReachObjective, [0.0, -0.94, 0.0, -2.8, 0.0, 1.8675, 0.0]),
    ("C2* panda reach (headline)", "config_panda_b200", 10000, PandaReachObjective, [0.0, -0.94, 0.0, -2.8, 0.0, 1.8675, 0.0]),
    ("C3 boxer_push", "config_boxer_push_b200", 4000, lambda: PushObjective(robot="boxer", link="ee_link"), [0.0, 2.5, 0.0]),
    ("C4 heijn_push (1 of 4 GPUs)", "config_heijn_push_b200", 4000, PushObjective, [0.0, 0.0, 0.0]),
    ("C4 heijn_push (all on 1 GPU)", "config_heijn_push_b200", 16000, PushObjective, [0.0, 0.0, 0.0]),
    ("C5 panda_pick (1 of 8 GPUs)", "config_panda_pick_b200", 8192, PandaPickObjective, [0.0, -0.94, 0.0, -2.8, 0.0, 1.8675, 0.0, 0.02, 0.02]),
    ("C5 panda_pick (all on 1 GPU)", "config_panda_pick_b200", 65536, PandaPickObjective, [0.0, -0.94, 0.0, -2.8, 0.0, 1.8675, 0.0, 0.02, 0.02]),
]

print("| config | K | T | nu | plan ms | plan Hz | rollout-steps/s | rollout us | cost us | reduce us |")
print("|---|---|---|---|---|---|---|---|---|---|")
for name, cfgname, K, obj, q in CASES:
    cfg = copy.deepcopy(load_isaacgym_config(cfgname))
    cfg.mppi.num_samples, cfg.mppi.device = K, "cuda:0"
    planner = MPPIisaacPlanner(cfg, obj(), use_cuda_graph=True)
    planner.sim.reset_robot_state(q, [0.0] * len(q))
    for _ in range(3):
        planner.mppi.command()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    a.record()
    for _ in range(n):
        planner.mppi.command()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    m = planner.mppi
    T, nu = m.T, m.nu
    kt = {}
    try:
        kt = bench.time_kernels(planner, reps=5)
    except Exception as e:  # noqa: BLE001
        kt = {"rollout_us": float("nan"), "cost_objective_torch_us": float("nan"), "reduce_us_warm_l2": float("nan")}
    print(f"| {name} | {K} | {T} | {nu} | {ms:.3f} | {1e3 / ms:.0f} | {K * T * 1e3 / ms:.3e} | {kt['rollout_us']:.0f} | "
          f"{kt['cost_objective_us']:.0f} | {kt['reduce_us_warm_l2']:.1f} |", flush=True)
    del planner
    torch.cuda.empty_cache()
