"""-m gpu: BOTH K2 mappings of trees and contact scenes -- one thread per rollout (rollout.cu + contact.cuh) and a team of lanes per
rollout (rollout_team.cu) -- against the oracle, whichever of the two the library would pick by itself for the scene and K
(mppib_rollout_mapping, include/mppib.h).  The parity tests of test_gpu_parity.py / test_gpu_sizes.py are re-run with the mapping forced
through MPPIB_K2_TEAM (read when a handle is created), with their own tolerances."""
import numpy as np
import pytest
import torch

import test_gpu_parity as P
import test_gpu_sizes as S
from scenes import boxer_setup, gripper_setup, panda_setup, push_setup

pytestmark = pytest.mark.gpu
MAPPINGS = [("thread", "0"), ("team", "1")]


@pytest.fixture(params=MAPPINGS, ids=[m[0] for m in MAPPINGS])
def mapping(request, monkeypatch):
    monkeypatch.setenv("MPPIB_K2_TEAM", request.param[1])
    return request.param[0]


def test_forced_mapping_is_what_runs(mapping):
    """the knob reaches the library: a tree (panda + gripper) reports the forced kernel, a serial chain stays on its lanes kernel"""
    sc, p, _ = gripper_setup(K=64, T=5)
    be = S.gpu_backend(sc, p)
    assert ("team" in be.rollout_mapping()) == (mapping == "team")
    sc, p, _ = panda_setup(K=64, T=5)
    assert "lanes" in S.gpu_backend(sc, p).rollout_mapping()


def test_default_mapping_by_scene(monkeypatch):
    """without the knob: serial chains -> lanes, every other scene the team kernel can take -> team, at every K (a shard of a multi-GPU
    job must run the same arithmetic as the single-GPU job)"""
    monkeypatch.delenv("MPPIB_K2_TEAM", raising=False)
    sc, p, _ = gripper_setup(K=64, T=5)
    assert "team" in S.gpu_backend(sc, p).rollout_mapping()
    for K in (4000, 65536):
        sc, p, _ = push_setup(K=K, T=5)
        assert "team" in S.gpu_backend(sc, p).rollout_mapping()
        sc, p, _ = S._pick_scene(K, 10)
        assert "team" in S.gpu_backend(sc, p).rollout_mapping()
    sc, p, _ = boxer_setup(K=4000, T=5)
    assert "team" in S.gpu_backend(sc, p).rollout_mapping()


def test_trees_free_running(oracle, mapping):
    P.test_k2_rollout_parity_free_running(oracle, gripper_setup, 512, 30)
    P.test_k2_drive_saturation_resolve_path(oracle, gripper_setup, 256)
    for actor, link, u_lim in [("albert", "mmrobot_link7", 0.4), ("omnipanda", "panda_hand", 0.3), ("jackal", "ee_link", 1.0)]:
        P.test_k2_rollout_parity_further_robots(oracle, actor, link, u_lim)


def test_contact_scenes_lockstep(oracle, mapping):
    P.test_contact_rollout_lockstep_parity(oracle)
    P.test_boxer_planar_base_rollout_parity(oracle)
    P.test_contact_randomisation_and_shard_offset_on_device(oracle)
    S.test_sphere_obstacle_rollout_lockstep(oracle)


def test_contact_scenes_at_baseline_sizes(oracle, mapping):
    S.test_c3_boxer_push_size_lockstep(oracle)
    S.test_c4_heijn_push_shard_size_lockstep(oracle)
    S.test_c5_panda_pick_shard_size_lockstep(oracle)


def test_contact_free_running_statistics(oracle, mapping):
    P.test_contact_rollout_free_running_statistics(oracle)


def test_contact_plans_through_planner_api(mapping):
    for task in ["push", "pick", "omnipick"]:
        P.test_contact_plan_through_planner_api(task)
    P.test_boxer_plan_through_planner_api()


def test_mappings_agree_on_one_step(oracle):
    """the two kernels from the SAME state, one model step of boxer_push at K = 512: float32 rounding apart (no oracle in between)"""
    import os
    K, T = 512, 8
    sc, p, s0 = boxer_setup(K=K, T=T)
    rng = np.random.default_rng(2)
    actions = np.stack([rng.uniform(0.3, 1.2, (T, K)), rng.uniform(-1.0, 1.0, (T, K))], axis=1).astype(np.float32)
    out = []
    for knob in ("0", "1"):
        os.environ["MPPIB_K2_TEAM"] = knob
        try:
            be = S.gpu_backend(sc, p)
        finally:
            del os.environ["MPPIB_K2_TEAM"]
        obs, state = torch.zeros((be.obs_size(), T, K), device=S.DEV), torch.zeros((be.state_size(), K), device=S.DEV)
        be.rollout(S.dev(s0), state, S.dev(actions), 0, 4, obs, root0=S.dev(sc.root_state0))
        out.append(state.cpu().numpy())
    assert np.isfinite(out[0]).all() and np.isfinite(out[1]).all()
    assert np.abs(out[0] - out[1]).max() < 5e-3
