#!/usr/bin/env python
"""Device time of the K2 rollout launch of a contact config at a given K under both mappings (thread-per-rollout / team of lanes):
    python tools/team_time.py pick 8192 [heijn 4000 ...]"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mppi_isaac_b200 import MPPIisaacPlanner, load_isaacgym_config  # noqa: E402
from mppi_isaac_b200.objectives import PandaPickObjective, PushObjective  # noqa: E402

SCENES = {
    "heijn": ("config_heijn_push_b200", PushObjective, [0.0] * 3),
    "boxer": ("config_boxer_push_b200", lambda: PushObjective(robot="boxer", link="ee_link"), [0.0, 2.5, 0.0]),
    "pick": ("config_panda_pick_b200", PandaPickObjective, [0.0, -0.94, 0.0, -2.8, 0.0, 1.8675, 0.0, 0.02, 0.02]),
}


def rollout_us(which, K, team, plans=4, reps=10):
    os.environ["MPPIB_K2_TEAM"] = "1" if team else "0"
    cfgname, obj, q = SCENES[which]
    cfg = copy.deepcopy(load_isaacgym_config(cfgname))
    cfg.mppi.num_samples, cfg.mppi.device = K, "cuda:0"
    planner = MPPIisaacPlanner(cfg, obj(), use_cuda_graph=False)
    planner.sim.reset_robot_state(q, [0.0] * len(q))
    for _ in range(plans):                      # a few plans so the sampled actions are the ones of a running controller
        planner.mppi.command()
    m = planner.mppi
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        planner.sim.rollout_all(m.actions)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def raw_us(actor, link, K, team, T=30, reps=10):
    """contact-free robots of conf/actors through the bare backend (tests/scenes.robot_setup), random commands"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    from scenes import robot_setup
    from mppi_isaac_b200.backend import CudaBackend
    os.environ["MPPIB_K2_TEAM"] = "1" if team else "0"
    sc, p, state0 = robot_setup(actor, link, K=K, T=T, u_lim=0.3)
    be = CudaBackend("cuda:0")
    be.create(sc.model, p)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).to("cuda:0")
    actions = dev(np.random.default_rng(8).uniform(-0.3, 0.3, (T, sc.nu, K)))
    obs, state = torch.zeros((be.obs_size(), T, K), device="cuda:0"), torch.zeros((be.state_size(), K), device="cuda:0")
    s0, root0 = dev(state0), dev(sc.root_state0.astype(np.float32))
    ts = []
    for _ in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        be.rollout(s0, state, actions, 0, T, obs, root0=root0)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:])
    return ts[len(ts) // 2], ts[0]


RAW = {"albert": "mmrobot_link7", "omnipanda": "panda_hand", "jackal": "ee_link", "boxer": "ee_link", "panda_gripper": "panda_hand", "heijn": "front_link"}

if __name__ == "__main__":
    args = sys.argv[1:]
    for which, K in zip(args[0::2], args[1::2]):
        if which.startswith("raw:"):
            a, b = raw_us(which[4:], RAW[which[4:]], int(K), False), raw_us(which[4:], RAW[which[4:]], int(K), True)
            print(f"{which:18s} K={int(K):6d}  thread {a[0]:9.1f} us (min {a[1]:9.1f})   team {b[0]:9.1f} us (min {b[1]:9.1f})   thread/team {a[0] / b[0]:.2f}", flush=True)
            continue
        a, b = rollout_us(which, int(K), False), rollout_us(which, int(K), True)
        print(f"{which:6s} K={int(K):6d}  thread {a[0]:9.1f} us (min {a[1]:9.1f})   team {b[0]:9.1f} us (min {b[1]:9.1f})   thread/team {a[0] / b[0]:.2f}", flush=True)
