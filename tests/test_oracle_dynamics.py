"""Pins the oracle's rollout (the spec the CUDA kernel is held to) with checks that do not share code with it."""
import copy

import numpy as np
import pytest

from mppi_isaac_b200.model.blob import OBS_DOF_STATE, OBS_LINK_STATE
from mppi_isaac_b200.model.urdf import forward_kinematics
from scenes import panda_setup, point_setup
from lagrange_ref import forward_dynamics


def _one_substep_qdd(oracle, sc, p, q, qd, u, use_double=True):
    """qdd of the oracle's first substep = (qd_new - qd)/h, read back through the DOF observation."""
    p = copy.copy(p)
    K, T, nu = p.K, p.T, sc.nu
    actions = np.zeros((T, nu, K), np.float32)
    actions[0] = np.asarray(u, np.float32)[:, None]
    state0 = np.concatenate([q, qd]).astype(np.float32)
    st, obs = oracle.rollout(sc.model, p, state0, actions, 0, 1, use_double=use_double)
    nb = sc.ndof
    h = p.dt / p.substeps
    assert p.substeps == 1
    qd_new = st[nb:2 * nb, 0].astype(np.float64)
    return (qd_new - np.asarray(qd, np.float64)) / h


@pytest.mark.parametrize("robot", ["panda", "point"])
def test_aba_matches_lagrangian(oracle, robot):
    from mppi_isaac_b200.utils.config_store import IsaacGymConfig
    rng = np.random.default_rng(3)
    if robot == "panda":
        sc, p, _ = panda_setup(K=4, T=2, sim=IsaacGymConfig(dt=0.001, substeps=1))
    else:
        sc, p, _ = point_setup(K=4, T=2)
        p.dt, p.substeps = 0.001, 1
    m = sc.model
    nb = sc.ndof
    for trial in range(4):
        lo = np.maximum(np.array(m.q_lo[:nb]), -2.5) + 0.2
        hi = np.minimum(np.array(m.q_hi[:nb]), 2.5) - 0.2
        q = rng.uniform(lo, hi)
        qd = rng.uniform(-0.8, 0.8, nb)
        u = rng.uniform(-0.2, 0.2, sc.nu)
        h = p.dt / p.substeps
        kd, b = np.array(m.kd[:nb], np.float64), np.array(m.damping[:nb], np.float64)
        target = np.array([m.cmd_c0[i] * u[m.cmd_i0[i]] + m.cmd_c1[i] * u[m.cmd_i1[i]] for i in range(nb)])
        tau = kd * (target - qd) - b * qd
        dimp = h * (kd + b)
        grav = (0.0, 0.0, -9.8) if m.gravity_on else (0.0, 0.0, 0.0)
        q32, qd32 = q.astype(np.float32).astype(np.float64), qd.astype(np.float32).astype(np.float64)
        qdd_ref, M = forward_dynamics(sc.robot, q32, qd32, tau, dimp, grav)
        # drive force limit: joints whose implicit drive torque exceeds URDF effort are re-solved once, saturated
        td = kd * (target - (qd + h * qdd_ref.numpy()))
        eff = np.array(m.effort[:nb], np.float64)
        sat = np.abs(td) > eff
        if sat.any():
            tau = np.where(sat, np.sign(td) * eff - b * qd, tau)
            dimp = np.where(sat, h * b, dimp)
            qdd_ref, M = forward_dynamics(sc.robot, q32, qd32, tau, dimp, grav)
        qdd = _one_substep_qdd(oracle, sc, p, q, qd, u)
        # float32 inputs / model constants limit agreement to ~1e-4 relative of |qdd| ~ 1e2
        np.testing.assert_allclose(qdd, qdd_ref.numpy(), rtol=2e-4, atol=2e-3)


def test_gravity_and_low_gain(oracle):
    """Same check with gravity on and a weak drive, so that M(q), Coriolis and gravity all matter."""
    rng = np.random.default_rng(5)
    sc, p, _ = panda_setup(K=4, T=2)
    p.dt, p.substeps = 0.001, 1
    m = copy.copy(sc.model)
    m.gravity_on = 1
    nb = sc.ndof
    for i in range(nb):
        m.kd[i], m.damping[i] = 0.5, 0.1
    sc.model = m
    for trial in range(3):
        q = rng.uniform(-1.5, 1.5, nb); q[3] = rng.uniform(-2.5, -0.5); q[5] = rng.uniform(0.5, 3.0)
        qd = rng.uniform(-1.0, 1.0, nb)
        u = rng.uniform(-0.2, 0.2, nb)
        h = 0.001
        tau = 0.5 * (u - qd) - 0.1 * qd
        qdd_ref, _ = forward_dynamics(sc.robot, q.astype(np.float32).astype(np.float64), qd.astype(np.float32).astype(np.float64), tau, np.full(nb, h * 0.6), (0, 0, -9.8))
        qdd = _one_substep_qdd(oracle, sc, p, q, qd, u)
        np.testing.assert_allclose(qdd, qdd_ref.numpy(), rtol=2e-4, atol=5e-3)


def test_velocity_drive_tracks_command(oracle):
    """Contact-free chain under the damping-600 velocity drive: q(T) ~= q0 + dt * sum(clamped u)."""
    sc, p, state0 = panda_setup(K=8, T=30)
    rng = np.random.default_rng(0)
    actions = rng.uniform(-0.2, 0.2, (30, 7, 8)).astype(np.float32)
    st, obs = oracle.rollout(sc.model, p, state0, actions)
    q_end = st[:7]
    expect = state0[:7, None] + p.dt * actions.sum(0)
    assert np.abs(q_end - expect).max() < 5e-3          # lag of order I/(kd) per step
    # DOF observation is interleaved q0,qd0,q1,qd1 (isaacgym_wrapper.py:190-192)
    dof = obs[13:, -1, :]
    np.testing.assert_allclose(dof[0::2], st[:7], atol=0)
    np.testing.assert_allclose(dof[1::2], st[7:], atol=0)


def test_observed_link_pose_is_fk_of_observed_q(oracle):
    sc, p, state0 = panda_setup(K=4, T=6, obs_links=("panda_ee_tip", "panda_link7", "panda_link0"))
    rng = np.random.default_rng(1)
    actions = rng.uniform(-0.2, 0.2, (6, 7, 4)).astype(np.float32)
    st, obs = oracle.rollout(sc.model, p, state0, actions, use_double=True)
    names = sc.robot.link_names
    for k in range(4):
        q = obs[39::2, -1, k][:7]
        pos, quat = forward_kinematics(sc.robot, q)
        for j, n in enumerate(("panda_ee_tip", "panda_link7", "panda_link0")):
            row = obs[13 * j:13 * j + 13, -1, k]
            np.testing.assert_allclose(row[0:3], pos[names.index(n)], atol=2e-6)
            qa, qb = row[3:7], quat[names.index(n)]
            assert min(np.abs(qa - qb).max(), np.abs(qa + qb).max()) < 2e-6
    # base link does not move
    np.testing.assert_allclose(obs[26:29], 0, atol=0)
    np.testing.assert_allclose(obs[33:39], 0, atol=0)


def test_link_velocity_is_time_derivative_of_position(oracle):
    sc, p, state0 = panda_setup(K=2, T=4)
    p.dt, p.substeps = 1e-3, 1
    actions = np.full((4, 7, 2), 0.15, np.float32)
    st, obs = oracle.rollout(sc.model, p, state0, actions, use_double=True)
    pos, vel = obs[0:3].astype(np.float64), obs[7:10].astype(np.float64)
    fd = (pos[:, 3, :] - pos[:, 2, :]) / p.dt
    # semi-implicit Euler: q3 = q2 + h*qd3, so the finite difference matches the NEW velocity to O(h)
    np.testing.assert_allclose(fd, vel[:, 3, :], atol=3e-4)


def test_replica_determinism(oracle):
    """Identical inputs in different sample slots evolve bit-identically
    (the one invariant of the reference's only test, mppiisaac/planner/tests/test_isaacgym_wrapper.py:35)."""
    sc, p, state0 = panda_setup(K=16, T=10)
    a = np.random.default_rng(2).uniform(-0.2, 0.2, (10, 7, 1)).astype(np.float32)
    actions = np.repeat(a, 16, axis=2)
    st, obs = oracle.rollout(sc.model, p, state0, actions)
    assert np.all(obs == obs[:, :, :1])
    assert np.all(st == st[:, :1])


def test_joint_limits_and_velocity_limits(oracle):
    sc, p, state0 = panda_setup(K=4, T=30)
    m = sc.model
    s0 = state0.copy()
    s0[3] = m.q_hi[3] - 0.01                      # joint 4 close to its upper limit (-0.0698)
    actions = np.zeros((30, 7, 4), np.float32); actions[:, 3, :] = 0.2
    st, obs = oracle.rollout(m, p, s0, actions)
    assert np.all(st[3] <= m.q_hi[3] + 1e-7) and np.all(st[3] >= m.q_hi[3] - 1e-4)
    assert np.all(np.abs(st[7 + 3]) < 1e-6)


def test_continue_from_state_equals_single_rollout(oracle):
    """Step-wise protocol (nsteps=1 from the stored state) == one T-step launch."""
    sc, p, state0 = panda_setup(K=8, T=6)
    actions = np.random.default_rng(4).uniform(-0.2, 0.2, (6, 7, 8)).astype(np.float32)
    st_all, obs_all = oracle.rollout(sc.model, p, state0, actions)
    state = np.repeat(state0[:, None], 8, 1).copy()
    obs = None
    for t in range(6):
        state, o = oracle.rollout(sc.model, p, None, actions, t, 1, state=state)
        obs = o if obs is None else np.where(np.arange(6)[None, :, None] == t, o, obs)
    np.testing.assert_array_equal(st_all, state)
    np.testing.assert_array_equal(obs_all, obs)


def test_diff_drive_command_map():
    """u=[0.2, 0], r=0.08, L=0.494 -> both wheels 2.5 rad/s (test_isaacgym_wrapper.py:27, test_boxer_config.yaml:8-9)."""
    r, L = 0.08, 0.494
    u = np.array([0.2, 0.0])
    left = u[0] / r - L * u[1] / (2 * r)
    right = u[0] / r + L * u[1] / (2 * r)
    assert abs(left - 2.5) < 1e-12 and abs(right - 2.5) < 1e-12
