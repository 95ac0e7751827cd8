#!/usr/bin/env python
"""World process of the panda reach example (role of the reference's examples/panda/world.py, without the viewer).

    python examples/panda/world.py [--connect tcp://127.0.0.1:4242] [--steps 200]

The "real" robot is a single-environment ``RolloutSim`` (the reference uses a one-env ``IsaacGymWrapper`` the same way); every control
step it ships its DOF and root state to the planner as ``torch.save`` bytes, applies the returned joint-velocity command and steps.
"""
import argparse
import copy
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

from mppi_isaac_b200 import load_isaacgym_config  # noqa: E402
from mppi_isaac_b200.planner.rollout_sim import RolloutSim  # noqa: E402
from mppi_isaac_b200.utils.rpc import RpcClient  # noqa: E402
from mppi_isaac_b200.utils.transport import bytes_to_torch, torch_to_bytes  # noqa: E402


def build_world(task="config_panda_b200", device=None, backend=None):
    cfg = copy.deepcopy(load_isaacgym_config(task))
    return cfg, RolloutSim(cfg.isaacgym, actors=cfg.actors, init_positions=cfg.initial_actor_positions, num_envs=1,
                           device=device or cfg.mppi.device, backend=backend, observe="all")


def control_step(sim, planner):
    """One closed-loop step: state -> planner -> command -> simulate (examples/panda/world.py:35-50 of the reference)."""
    action = bytes_to_torch(planner.compute_action_tensor(torch_to_bytes(sim._dof_state.cpu()), torch_to_bytes(sim._root_state.cpu())))
    sim.apply_robot_cmd(action.to(sim.device))
    sim.step()
    return action


def goal_distance(sim, link="panda_ee_tip", actor="panda"):
    ee = sim.get_actor_link_by_name(actor, link)[0, 0:3]
    return float(torch.linalg.norm(ee - sim.get_actor_position_by_name("goal")[0, 0:3]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="config_panda_b200")
    ap.add_argument("--connect", default="tcp://127.0.0.1:4242")
    ap.add_argument("--steps", type=int, default=200)
    args = ap.parse_args()
    cfg, sim = build_world(args.task)
    planner = RpcClient(args.connect)
    print("Mppi server found!" if planner.compute_action_tensor(torch_to_bytes(sim._dof_state.cpu()), torch_to_bytes(sim._root_state.cpu())) else "")
    t = time.time()
    for i in range(args.steps):
        control_step(sim, planner)
        dt = time.time() - t
        t = time.time()
        if i % 10 == 0:
            print(f"step {i:4d}  |ee - goal| = {goal_distance(sim):.3f} m   loop {1.0 / max(dt, 1e-9):7.1f} Hz  (real time x{cfg.isaacgym.dt / max(dt, 1e-9):.1f})")


if __name__ == "__main__":
    main()
