"""Plugin fixtures: the example Objectives of the BASELINE configs, restated.

The reference ships its cost functions inside ``examples/*/planner.py``; they are *user* plugins of the
planner (protocol: ``compute_cost(sim) -> (N,) tensor``, ``reset()``, optional ``weights``), not part of
the library.  They are restated here so the parity tests and ``bench.py`` can drive the drop-in surface
with the same arithmetic:

* ``PandaReachObjective``   <- ``examples/panda/planner.py:22-40``      (O1, config C2)
* ``PointReachObjective``   <- ``examples/heijn_reach/planner.py:15-24`` (O4 template, config C1)

Every term is row-wise over dim 0, so the functions work unchanged on the (K, .) step views and the
(T*K, .) batched views of ``RolloutSim``.
"""
import torch

from .utils.conversions import matrix_to_euler_angles, quaternion_to_matrix


class PandaReachObjective:
    """w_goal * |p_ee - p_goal| + w_ori * |euler_ZYX(R(q_ee))[:2]| for the stick tip."""

    def __init__(self, cfg=None, actor: str = "panda", link: str = "panda_ee_tip", goal: str = "goal"):
        self.weights = {"robot_to_goal": 1.0, "robot_ori": 0.5}
        self.actor, self.link, self.goal = actor, link, goal

    def reset(self):
        pass

    def compute_cost(self, sim):
        ee = sim.get_actor_link_by_name(self.actor, self.link)
        goal = sim.get_actor_position_by_name(self.goal)
        dist = torch.linalg.norm(ee[:, 0:3] - goal[:, 0:3], axis=1)
        # the reference hands the xyzw quaternion to a real-first API; reproduced literally (Appendix A #11)
        zyx = matrix_to_euler_angles(quaternion_to_matrix(ee[:, 3:7]), "ZYX")[:, 0:2]
        ori = torch.linalg.norm(zyx, axis=1)
        return self.weights["robot_to_goal"] * dist + self.weights["robot_ori"] * ori


class PointReachObjective:
    """Planar distance of a robot link to the goal actor (+ optional wall contact force)."""

    def __init__(self, cfg=None, actor: str = "point_robot", link: str = "base_link", goal: str = "goal", wall: str = None):
        self.actor, self.link, self.goal, self.wall = actor, link, goal, wall
        self.weights = {}

    def reset(self):
        pass

    def compute_cost(self, sim):
        r = sim.get_actor_link_by_name(actor_name=self.actor, link_name=self.link)
        g = sim.get_actor_position_by_name(self.goal)
        cost = torch.linalg.norm(g[:, 0:2] - r[:, 0:2], axis=1)
        if self.wall is not None:
            f = sim.get_actor_contact_forces_by_name(self.wall, "box")
            cost = cost + torch.sum(torch.abs(f[:, 0:3]), axis=1)
        return cost
