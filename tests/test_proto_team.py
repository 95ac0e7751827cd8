"""CPU (no GPU): the formulation of the K2 team kernel's articulation phase (csrc/rollout_team.cu) -- frames and spatial velocities by
pointer jumping over the ancestors, composite rigid bodies as differences of suffix sums over the depth-first body order, joint-space
LDL^T with the leaves eliminated first -- restated in float64 numpy (tests/proto_team.py) and checked against the oracle's body-frame
articulated-body algorithm on the tree robots of conf/actors, incl. a rollout driven far beyond the effort limits (saturation
re-solve) and the planar differential-drive bases.  Also pinned here: the pivots of the leaves-first elimination ARE the
articulated-body diagonals D_j, which the contact solve uses as the joints' compliance."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import proto_team  # noqa: E402


@pytest.mark.parametrize("robot", ["panda_gripper", "omnipanda", "albert", "jackal", "boxer"])
def test_tree_formulation_matches_the_oracle(oracle, robot):
    assert proto_team.check(names=(robot,), K=2, T=4, verbose=False) < 5e-5


def test_tree_tables_of_a_branching_robot():
    """panda + gripper: 7 arm bodies in a chain, two fingers on the hand (body 6)"""
    parent = [-1, 0, 1, 2, 3, 4, 5, 6, 6]
    jump, desc, end = proto_team.tree_tables(parent, 16)
    assert jump[0] == parent and jump[1][8] == 5 and jump[2][8] == 3 and jump[3][8] == -1
    assert desc[6] == {6, 7, 8} and desc[7] == {7} and end == [9, 9, 9, 9, 9, 9, 9, 8, 9]
    x = [np.array([float(i + 1)]) for i in range(9)]
    assert [float(v[0]) for v in proto_team.anc_sum(x, jump)] == [1, 3, 6, 10, 15, 21, 28, 36, 37]
    assert [float(v[0]) for v in proto_team.subtree_sum(x, end)] == [45, 44, 42, 39, 35, 30, 24, 8, 9]
    with pytest.raises(AssertionError):
        proto_team.tree_tables([-1, 0, 0, 1], 8)          # body 3 hangs below body 1 but is numbered after body 2: not depth first


def test_generalised_coordinate_contact_solve_matches_the_oracle(oracle):
    """a tilted, spinning free box on the ground, one model step: Gauss-Seidel over [v (world); omega (body axes)] with one scalar inverse
    inertia per coordinate and rows cached per contact (the team kernel's formulation) == the oracle's world-frame solve"""
    assert proto_team.check_free_box(verbose=False) < 5e-6


@pytest.mark.parametrize("scene,steps,K", [("heijn", (0, 3, 6, 9), 48), ("boxer", (0, 4, 7, 9), 48), ("pick", (0, 10, 20, 29), 12)])
def test_whole_contact_step_matches_the_oracle(oracle, scene, steps, K):
    """the team kernel's whole step restated in float64 -- tree articulation, contacts in the oracle's order up to the contact
    capacity, Gauss-Seidel over joints + free-body components -- in lock-step with the oracle on the C4 / C3 push scenes while the
    robot pushes the block (up to 24 contacts per substep), and on panda_pick (9-joint tree, block on the table)"""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    wx, wv = proto_team.check_push_scene(scene, verbose=False, steps=steps, K=K)
    assert wx < 2e-6 and wv < 2e-5


def test_lanes_formulation_matches_the_oracle(oracle):
    """the serial-chain special case (csrc/rollout_lanes.cu: Kogge-Stone scans, distributed LDL^T from the root) -- tests/proto_lanes.py"""
    import proto_lanes
    proto_lanes.main()
