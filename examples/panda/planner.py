#!/usr/bin/env python
"""Planner process of the panda reach example (role of the reference's examples/panda/planner.py).

    python examples/panda/planner.py [--task config_panda_b200] [--bind tcp://0.0.0.0:4242] [--samples K]

Serves ``MPPIisaacPlanner`` over the zerorpc-compatible request/reply protocol of ``mppi_isaac_b200.utils.rpc``: the world process
(``world.py``, or the reference's own ``world.py`` through ``zerorpc.Client``) calls ``compute_action_tensor(dof_bytes, root_bytes)``
once per control step.  The Objective below is user code, exactly as in the reference: any ``compute_cost(sim) -> (N,) tensor``.
"""
import argparse
import copy
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from mppi_isaac_b200 import MPPIisaacPlanner, load_isaacgym_config, ops  # noqa: E402
from mppi_isaac_b200.utils.rpc import RpcServer  # noqa: E402


class Objective(object):
    """Reach a goal with the stick tip and keep it upright: w_goal |p_ee - p_goal| + w_ori |euler_ZYX(ee)[:2]|."""

    def __init__(self, cfg=None):
        self.weights = {"robot_to_goal": 1.0, "robot_ori": 0.5}

    def reset(self):
        pass

    def compute_cost(self, sim):
        ee = sim.get_actor_link_by_name("panda", "panda_ee_tip")          # (N, 13): position, quaternion xyzw, velocities
        goal = sim.get_actor_position_by_name("goal")                      # (N, 3)
        return ops.pose_cost(ee, goal, self.weights["robot_to_goal"], self.weights["robot_ori"])   # one fused kernel (torch ops work too)


def build_planner(task="config_panda_b200", samples=None, device=None, backend=None):
    cfg = copy.deepcopy(load_isaacgym_config(task))
    if samples:
        cfg.mppi.num_samples = int(samples)
    if device:
        cfg.mppi.device = device
    return MPPIisaacPlanner(cfg, Objective(cfg), prior=None, backend=backend)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="config_panda_b200")
    ap.add_argument("--bind", default="tcp://0.0.0.0:4242")
    ap.add_argument("--samples", type=int, default=None)
    args = ap.parse_args()
    server = RpcServer(build_planner(args.task, args.samples)).bind(args.bind)
    print(f"MPPI planner serving on {args.bind}")
    server.run()


if __name__ == "__main__":
    main()
