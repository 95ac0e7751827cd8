#!/usr/bin/env python
"""Where do the host-side microseconds of one end-to-end plan go?   python tools/e2e_breakdown.py [iters]
Times the stages of MPPIisaacPlanner.compute_action_tensor (bytes in -> bytes out) and compute_action (q, qdot -> tensor)
with a synchronize after every stage (so the stage sums exceed the pipelined totals printed first)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from mppi_isaac_b200 import MPPIisaacPlanner  # noqa: E402
from mppi_isaac_b200.objectives import PandaReachObjective  # noqa: E402
from mppi_isaac_b200.utils.transport import bytes_to_torch, torch_to_bytes  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
planner = MPPIisaacPlanner(bench.panda_cfg(bench.K_PER_GPU, "cuda:0"), PandaReachObjective(), use_cuda_graph=True)
q0, goal = bench.synthetic_state()
dof_b, root_b, _ = bench.world_bytes(planner, q0, np.zeros(7), goal)
for _ in range(5):
    planner.compute_action_tensor(dof_b, root_b)
torch.cuda.synchronize()


def timeit(fn, n=N):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


rng = np.random.default_rng(3)
inputs = [bench.world_bytes(planner, q0 + rng.uniform(-0.05, 0.05, 7), rng.uniform(-0.1, 0.1, 7), goal)[:2] for _ in range(N)]
it = iter(range(10 ** 9))
print(f"compute_action_tensor (bytes -> bytes)   {timeit(lambda: planner.compute_action_tensor(*inputs[next(it) % N])):8.1f} us")
qs = [list(q0 + rng.uniform(-0.05, 0.05, 7)) for _ in range(N)]
zero = [0.0] * 7
for _ in range(5):
    planner.compute_action(qs[0], zero)
print(f"compute_action (q, qdot -> cpu tensor)   {timeit(lambda: planner.compute_action(qs[next(it) % N], zero)):8.1f} us")
print(f"mppi.command() device only (graph)       {timeit(lambda: planner.mppi.command()):8.1f} us")


def sync(fn):
    def g():
        r = fn()
        torch.cuda.synchronize()
        return r
    return g


dof_t, root_t = bytes_to_torch(dof_b), bytes_to_torch(root_b)
act = planner.mppi.command()
print("-- stages of compute_action_tensor, each followed by a synchronize")
print(f"objective.reset                          {timeit(sync(planner.objective.reset)):8.1f} us")
print(f"bytes_to_torch(dof)                      {timeit(lambda: bytes_to_torch(dof_b)):8.1f} us")
print(f"bytes_to_torch(root)                     {timeit(lambda: bytes_to_torch(root_b)):8.1f} us")
print(f"sim.set_world_state(dof, root)           {timeit(sync(lambda: planner.sim.set_world_state(dof_t, root_t))):8.1f} us")
print(f"sim.set_world_state(dof, None)           {timeit(sync(lambda: planner.sim.set_world_state(dof_t, None))):8.1f} us")
print(f"mppi.command() + sync                    {timeit(sync(planner.mppi.command)):8.1f} us")
print(f"torch_to_bytes(cuda action)              {timeit(lambda: torch_to_bytes(act)):8.1f} us")
print(f"action.cpu()                             {timeit(lambda: act.cpu()):8.1f} us")
cpu_act = act.cpu()
print(f"torch_to_bytes(cpu action)               {timeit(lambda: torch_to_bytes(cpu_act)):8.1f} us")
print("-- stages of compute_action")
print(f"sim.reset_root_state                     {timeit(sync(planner.sim.reset_root_state)):8.1f} us")
print(f"sim.reset_robot_state(q, qdot)           {timeit(sync(lambda: planner.sim.reset_robot_state(qs[0], zero))):8.1f} us")
print(f"sim.save_root_state                      {timeit(sync(planner.sim.save_root_state)):8.1f} us")
