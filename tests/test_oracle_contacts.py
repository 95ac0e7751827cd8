"""Physics sanity of the oracle's free-body / contact restatement (the spec of DESIGN.md section 2 for PhysX rigid
bodies): analytic rest / friction answers, pushing, blocking by static boxes, per-rollout randomisation."""
import numpy as np
import pytest

from scenes import push_setup

G = 9.8
ROW_FRONT, ROW_DOF, ROW_BLOCK, ROW_FBLOCK = 0, 13, 19, 32       # obs rows of push_setup


def run(oracle, sc, p, state0, actions, **kw):
    return oracle.rollout(sc.model, p, state0, actions, root0=sc.root_state0, **kw)


def test_block_rests_on_the_ground(oracle):
    sc, p, s0 = push_setup(K=4, T=20, block_pos=(3.0, 3.0, 0.1), obstacles=False)
    actions = np.zeros((20, 3, 4), np.float32)
    st, obs = run(oracle, sc, p, s0, actions, use_double=True)
    blk = obs[ROW_BLOCK:ROW_BLOCK + 13]
    np.testing.assert_allclose(blk[0:2, -1], [[3.0] * 4, [3.0] * 4], atol=2e-3)          # no horizontal drift
    sink = 0.1 - blk[2, -1, 0]
    assert -1e-4 <= sink < 2e-3                                                               # implicit penalty: ~ m g / kp
    np.testing.assert_allclose(blk[7:13, -1], 0, atol=2e-3)                               # at rest (8 Gauss-Seidel sweeps leave mm/s jitter)
    np.testing.assert_allclose(obs[ROW_FBLOCK:ROW_FBLOCK + 3, -1, 0], [0, 0, 1.0 * G], atol=2e-2)   # net contact force = weight


def test_sliding_block_decelerates_with_coulomb_friction(oracle):
    sc, p, s0 = push_setup(K=4, T=10, block_pos=(3.0, 3.0, 0.1), obstacles=False, dt=0.02, substeps=1)
    root0 = sc.root_state0.copy()
    root0[1, 7] = 2.0                                                                     # initial v_x = 2 m/s
    mu = 0.5 * (0.6 + 1.0)                                                                # average combine with the ground (mu 1)
    st, obs = oracle.rollout(sc.model, p, s0, np.zeros((10, 3, 4), np.float32), root0=root0, use_double=True)
    vx = obs[ROW_BLOCK + 7, :, 0]
    expect = 2.0 - mu * G * 0.02 * np.arange(1, 11)
    # the block starts exactly touching (no penetration, no contact in the very first step), then settles
    np.testing.assert_allclose(vx[3:], expect[3:], atol=0.03)
    np.testing.assert_allclose(np.diff(vx[5:]), -mu * G * 0.02, atol=2e-3)     # converges to mu g h per step
    # ... and comes to rest instead of reversing
    sc, p, s0 = push_setup(K=4, T=40, block_pos=(3.0, 3.0, 0.1), obstacles=False, dt=0.02, substeps=1)
    st, obs = oracle.rollout(sc.model, p, s0, np.zeros((40, 3, 4), np.float32), root0=root0, use_double=True)
    assert abs(obs[ROW_BLOCK + 7, -1, 0]) < 1e-3 and obs[ROW_BLOCK + 7].min() > -1e-3


def test_robot_pushes_block(oracle):
    sc, p, s0 = push_setup(K=4, T=25, block_pos=(0.7, 1.5, 0.1), obstacles=False)         # block 19.5 cm ahead of the base
    actions = np.zeros((25, 3, 4), np.float32); actions[:, 0] = 0.5                      # drive +x at 0.5 m/s
    st, obs = run(oracle, sc, p, s0, actions, use_double=True)
    bx, rx = obs[ROW_BLOCK, :, 0], obs[ROW_DOF, :, 0]                                     # block x, robot q_x
    assert rx[-1] > 1.0                                                                   # the drive keeps tracking the command
    gap = bx - (rx + 0.305 + 0.2)                                                         # block face minus robot front face
    assert gap[-1] > -0.03 and bx[-1] > 0.7 + 0.7                                         # pushed along, no tunnelling
    assert abs(obs[ROW_BLOCK + 7, -1, 0] - 0.5) < 0.1                                     # block moves with the robot
    assert abs(obs[ROW_FBLOCK + 2, 5:, 0].mean() - G) < 0.5                               # net contact force: weight carried, push ~ friction
    assert abs(obs[ROW_BLOCK + 1, -1, 0] - 1.5) < 0.1                                     # pushed roughly straight (Gauss-Seidel order breaks the symmetry a little)


def test_block_pushed_into_static_obstacle_transmits_force(oracle):
    # obstacle 1 spans x in [0.7,1.3], y in [0.6,1.4], only 10.8 cm high; push the 20 cm high block along +x into its -x face
    sc, p, s0 = push_setup(K=4, T=40, block_pos=(0.25, 1.0, 0.1), robot_pos=(-0.5, 1.0, 0.05))
    sc.root_state0[3, 0:3] = [5.0, 5.0, 0.054]                                            # park obstacle 2 far away
    actions = np.zeros((40, 3, 4), np.float32); actions[:, 0] = 0.4
    st, obs = run(oracle, sc, p, s0, actions, use_double=True)
    assert np.all(np.isfinite(obs))
    f_obst = obs[ROW_FBLOCK + 3:ROW_FBLOCK + 6]                                           # net force on paper_obst1
    assert f_obst[0].max() > 30.0                                                         # the push is transmitted into it (+x): the cost sees it
    np.testing.assert_allclose(obs[ROW_FBLOCK + 6:ROW_FBLOCK + 9], 0, atol=1e-9)          # obstacle 2 untouched
    t_hit = int(np.argmax(f_obst[0, :, 0] > 1.0))
    assert abs(obs[ROW_BLOCK, t_hit, 0] - (0.7 - 0.2)) < 0.06                             # first contact when the faces meet
    assert obs[ROW_DOF + 1, t_hit + 3:, 0].mean() < 0.3                                   # the drive is held back by the jam (commanded 0.4 m/s)


def test_robot_vs_static_obstacle_stops_the_robot(oracle):
    sc, p, s0 = push_setup(K=4, T=40, block_pos=(4.0, 4.0, 0.1), robot_pos=(0.2, 1.0, 0.05))
    sc.root_state0[3, 0:3] = [5.0, 5.0, 0.054]
    actions = np.zeros((40, 3, 4), np.float32); actions[:, 0] = 0.6
    st, obs = run(oracle, sc, p, s0, actions, use_double=True)
    rx = 0.2 + obs[ROW_DOF, :, 0]
    assert rx.max() < 0.7 - 0.305 + 0.05                                                  # front face is held at the obstacle face
    assert abs(obs[ROW_DOF + 1, -5:, 0]).max() < 0.05                                     # stalled
    f = obs[ROW_FBLOCK + 3]
    assert 150 < f[-5:].mean() < 600 and np.isfinite(obs).all()                           # ~ stall force of the velocity drive, 600 * 0.6 N


def test_per_rollout_randomisation_is_seeded_and_shard_invariant(oracle):
    sc, p, s0 = push_setup(K=64, T=2, noise=True, block_pos=(3.0, 3.0, 0.1), obstacles=False, dt=0.05)
    root0 = sc.root_state0.copy(); root0[1, 7] = 1.0
    a = np.zeros((2, 3, 64), np.float32)
    _, o1 = oracle.rollout(sc.model, p, s0, a, root0=root0)
    _, o2 = oracle.rollout(sc.model, p, s0, a, root0=root0)
    np.testing.assert_array_equal(o1, o2)                                                 # reproducible (unlike np.random in the reference)
    vx = o1[ROW_BLOCK + 7, -1]
    assert vx.std() > 1e-3                                                                # friction differs per rollout ...
    mu = (1.0 - vx) / (G * 0.05 * 2)
    assert 0.5 * (0.6 * 0.7 + 1) - 1e-3 <= mu.min() and mu.max() <= 0.5 * (0.6 * 1.3 + 1) + 1e-3   # ... within +-30 %
    p.K, p.k_offset = 32, 32                                                              # second shard of 2
    _, o3 = oracle.rollout(sc.model, p, s0, a[:, :, :32], root0=root0)
    np.testing.assert_array_equal(o3[:, :, :], o1[:, :, 32:])


def test_stepwise_continue_equals_batched_with_contacts(oracle):
    sc, p, s0 = push_setup(K=8, T=6, block_pos=(0.7, 1.5, 0.1))
    rng = np.random.default_rng(0)
    actions = rng.uniform(-0.6, 0.6, (6, 3, 8)).astype(np.float32); actions[:, 0] = 0.5
    st_all, obs_all = run(oracle, sc, p, s0, actions)
    NS = 6 + 13
    state = np.zeros((NS, 8), np.float32)
    state[6:19] = sc.root_state0[1][:, None]
    obs = np.zeros_like(obs_all)
    for t in range(6):
        state, o = oracle.rollout(sc.model, p, None, actions, t, 1, state=state, root0=sc.root_state0)
        obs[:, t] = o[:, t]
    np.testing.assert_array_equal(st_all, state)
    # the contact-force rows are per-substep quantities and identical too
    np.testing.assert_array_equal(obs_all, obs)
