"""Shared builders for the test scenes (BASELINE configs C1 / C2 on the built-in conf + compiled models)."""
import copy

import numpy as np

from mppi_isaac_b200.model.blob import (OBS_CONTACT, OBS_DOF_STATE, OBS_FREE_STATE, OBS_LINK_STATE, build_scene, make_params)
from mppi_isaac_b200.utils.config_store import IsaacGymConfig, MPPIConfig, load_actor_cfgs, load_isaacgym_config


def panda_scene():
    return build_scene(load_actor_cfgs(["panda_stick", "goal"]))


def point_scene():
    return build_scene(load_actor_cfgs(["point_robot", "goal"]))


def panda_mppi(K=64, T=30, mode="simple", **kw):
    d = dict(num_samples=K, horizon=T, mppi_mode=mode, sampling_method="random", noise_sigma=(0.1 * np.eye(7)).tolist(),
             u_min=[-0.2], u_max=[0.2], lambda_=0.05, sample_null_action=True, rollout_var_discount=0.95)
    d.update(kw)
    return MPPIConfig(**d)


def point_mppi(K=128, T=12, mode="simple", **kw):
    d = dict(num_samples=K, horizon=T, mppi_mode=mode, sampling_method="random", noise_sigma=np.eye(3).tolist(),
             u_min=[-1.5], u_max=[1.5], lambda_=0.1, sample_null_action=True, rollout_var_discount=0.95)
    d.update(kw)
    return MPPIConfig(**d)


def panda_setup(K=64, T=30, mode="simple", obs_links=("panda_ee_tip",), dof=True, sim=None, **kw):
    sc = panda_scene()
    obs = [(OBS_LINK_STATE, sc.robot.link_names.index(n)) for n in obs_links]
    if dof:
        obs.append((OBS_DOF_STATE, 0))
    p = make_params(panda_mppi(K, T, mode, **kw), sim or IsaacGymConfig(), sc.nu, K, obs)
    dof0 = sc.dof_state0
    state0 = np.concatenate([dof0[0::2], dof0[1::2]]).astype(np.float32)
    return sc, p, state0


def point_setup(K=128, T=12, mode="simple", **kw):
    sc = point_scene()
    obs = [(OBS_LINK_STATE, sc.robot.link_names.index("base_link")), (OBS_DOF_STATE, 0)]
    p = make_params(point_mppi(K, T, mode, **kw), IsaacGymConfig(), sc.nu, K, obs)
    state0 = np.array([0.1, 0, 0, 0, 0, 0], np.float32)
    return sc, p, state0


def gripper_setup(K=64, T=30, mode="simple", **kw):
    """panda + two-finger gripper: a TREE (both fingers hang off link 7) -> exercises the general-topology kernel path."""
    sc = build_scene(load_actor_cfgs(["panda_gripper", "goal"]))
    names = sc.robot.link_names
    obs = [(OBS_LINK_STATE, names.index("panda_ee")), (OBS_DOF_STATE, 0), (OBS_LINK_STATE, names.index("panda_leftfinger")),
           (OBS_LINK_STATE, names.index("panda_rightfinger"))]
    mc = panda_mppi(K, T, mode, noise_sigma=(0.1 * np.eye(9)).tolist(), **kw)
    p = make_params(mc, IsaacGymConfig(), sc.nu, K, obs)
    dof0 = sc.dof_state0
    state0 = np.concatenate([dof0[0::2], dof0[1::2]]).astype(np.float32)
    return sc, p, state0


def push_setup(K=32, T=10, noise=False, block_pos=(1.0, 1.5, 0.1), robot_pos=(0.0, 1.5, 0.05), dt=0.1, substeps=1, obstacles=True, **kw):
    """heijn omni base + pushable block (+ two static boxes): BASELINE config C4 geometry, contact path."""
    names = ["heijn", "block"] + (["paper_obst1", "paper_obst2"] if obstacles else []) + ["goal"]
    actors = load_actor_cfgs(names)
    actors[0].init_pos = list(robot_pos)
    actors[1].init_pos = list(block_pos)
    if not noise:
        for a in actors:
            a.noise_sigma_size, a.noise_percentage_mass, a.noise_percentage_friction = None, 0.0, 0.0
    sim = IsaacGymConfig(dt=dt, substeps=substeps)
    sc = build_scene(actors, substep=dt / substeps)
    rn = sc.robot.link_names
    obs = [(OBS_LINK_STATE, rn.index("front_link")), (OBS_DOF_STATE, 0), (OBS_FREE_STATE, 0), (OBS_CONTACT, sc.contact_slot[sc.body_offset[1]])]
    if obstacles:
        obs += [(OBS_CONTACT, sc.contact_slot[sc.body_offset[2]]), (OBS_CONTACT, sc.contact_slot[sc.body_offset[3]])]
    mc = MPPIConfig(num_samples=K, horizon=T, mppi_mode="simple", sampling_method="random", noise_sigma=[[0.5, 0, 0], [0, 0.5, 0], [0, 0, 1.8]],
                    u_min=[-0.6, -0.6, -1.0], u_max=[0.6, 0.6, 1.0], lambda_=0.05, sample_null_action=True, **kw)
    p = make_params(mc, sim, sc.nu, K, obs)
    state0 = np.zeros(6, np.float32)
    return sc, p, state0


def panda_cfg(K=64, T=30, device="cpu", **mppi_kw):
    cfg = load_isaacgym_config("config_panda_b200")
    cfg = copy.deepcopy(cfg)
    cfg.mppi.num_samples, cfg.mppi.horizon, cfg.mppi.device = K, T, device
    for k, v in mppi_kw.items():
        setattr(cfg.mppi, k, v)
    return cfg


def push_cfg(K=64, T=10, device="cpu", **mppi_kw):
    """BASELINE config C4 (heijn_push) at a test-sized K / T."""
    cfg = copy.deepcopy(load_isaacgym_config("config_heijn_push_b200"))
    cfg.mppi.num_samples, cfg.mppi.horizon, cfg.mppi.device = K, T, device
    for k, v in mppi_kw.items():
        setattr(cfg.mppi, k, v)
    return cfg


def pick_cfg(K=32, T=9, device="cpu", **mppi_kw):
    """BASELINE config C5 (panda_pick) at a test-sized K / T."""
    cfg = copy.deepcopy(load_isaacgym_config("config_panda_pick_b200"))
    cfg.mppi.num_samples, cfg.mppi.horizon, cfg.mppi.device = K, T, device
    for k, v in mppi_kw.items():
        setattr(cfg.mppi, k, v)
    return cfg


def boxer_cfg(K=64, T=12, device="cpu", **mppi_kw):
    """BASELINE config C3 (boxer_push: differential-drive base reduced to the plane) at a test-sized K / T."""
    cfg = copy.deepcopy(load_isaacgym_config("config_boxer_push_b200"))
    cfg.mppi.num_samples, cfg.mppi.horizon, cfg.mppi.device = K, T, device
    for k, v in mppi_kw.items():
        setattr(cfg.mppi, k, v)
    return cfg


def boxer_setup(K=32, T=10, noise=True):
    actors = load_actor_cfgs(["boxer", "block", "paper_obst1", "paper_obst2", "goal"])
    actors[0].init_pos = [0.0, 2.5, 0.05]
    actors[1].init_pos = [0.0, 1.75, 0.1]                  # block right in front of the chassis (the base drives along -y)
    if not noise:
        for a in actors:
            a.noise_sigma_size, a.noise_percentage_mass, a.noise_percentage_friction = None, 0.0, 0.0
    sim = IsaacGymConfig(dt=0.05, substeps=2)
    sc = build_scene(actors, substep=0.025)
    rn = sc.robot.link_names
    obs = [(OBS_LINK_STATE, rn.index("ee_link")), (OBS_LINK_STATE, 0), (OBS_DOF_STATE, 0), (OBS_FREE_STATE, 0),
           (OBS_CONTACT, sc.contact_slot[sc.body_offset[2]])]
    mc = MPPIConfig(num_samples=K, horizon=T, mppi_mode="simple", sampling_method="random", noise_sigma=[[2.0, 0], [0, 8.0]],
                    u_min=[-1.2, -3.5], u_max=[1.2, 3.5], lambda_=0.01, sample_null_action=True)
    p = make_params(mc, sim, sc.nu, K, obs)
    dof0 = sc.dof_state0
    state0 = np.concatenate([dof0[0::2], dof0[1::2]]).astype(np.float32)
    return sc, p, state0


def point_cfg(K=128, T=12, device="cpu", **mppi_kw):
    cfg = copy.deepcopy(load_isaacgym_config("config_point_robot_b200"))
    cfg.mppi.num_samples, cfg.mppi.horizon, cfg.mppi.device = K, T, device
    for k, v in mppi_kw.items():
        setattr(cfg.mppi, k, v)
    return cfg


def robot_setup(actor, link, K=64, T=10, u_lim=0.2, sigma=0.1, extra_actors=("goal",), substeps=2, dt=0.05, **kw):
    """Any single-robot scene of the shipped conf/ (albert, omnipanda, panda_effort, ...): link + DOF observations, contact-free unless
    `extra_actors` brings colliding boxes."""
    actors = load_actor_cfgs([actor] + list(extra_actors))
    for a in actors:
        a.noise_sigma_size, a.noise_percentage_mass, a.noise_percentage_friction = None, 0.0, 0.0
    sc = build_scene(actors, substep=dt / substeps)
    obs = [(OBS_LINK_STATE, sc.robot.link_names.index(link)), (OBS_DOF_STATE, 0)]
    if sc.model.nfree:
        obs.append((OBS_FREE_STATE, 0))
    mc = MPPIConfig(num_samples=K, horizon=T, mppi_mode="simple", sampling_method="random", noise_sigma=(sigma * np.eye(sc.nu)).tolist(),
                    u_min=[-u_lim], u_max=[u_lim], lambda_=0.05, sample_null_action=True, **kw)
    p = make_params(mc, IsaacGymConfig(dt=dt, substeps=substeps), sc.nu, K, obs)
    dof0 = sc.dof_state0                                  # interleaved (q, qd) of ALL joints, virtual base joints included
    state0 = np.concatenate([dof0[0::2], dof0[1::2]]).astype(np.float32)
    return sc, p, state0
