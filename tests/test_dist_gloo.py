"""Sample sharding over ranks (N>1 path) on CPU: world_size-2 gloo processes, checker backend.
The sharded plan (each rank rolls out its K/G samples, one all-gather of the (beta, eta, W) partials, K4 on every
rank) must reproduce the single-process plan, and every rank must end with the same U."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
Q0 = [0.0, -0.94, 0.0, -2.8, 0.0, 1.8675, 0.0]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mppi_isaac_b200 import MPPIisaacPlanner
    from mppi_isaac_b200.objectives import PandaReachObjective
    from oracle.backend import OracleBackend
    from scenes import panda_cfg
    p = MPPIisaacPlanner(panda_cfg(K=64, T=12), PandaReachObjective(), backend=OracleBackend())
    assert p.sim.num_envs == 64 // world and p.k_offset == rank * (64 // world)
    acts = [p.compute_action(Q0, [0] * 7).numpy() for _ in range(3)]
    np.save(os.path.join(out_dir, f"act_{rank}.npy"), np.stack(acts))
    np.save(os.path.join(out_dir, f"U_{rank}.npy"), p.mppi.U.numpy())
    np.save(os.path.join(out_dir, f"actions_{rank}.npy"), p.mppi.actions.numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_plan_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, HERE)
    from mppi_isaac_b200 import MPPIisaacPlanner
    from mppi_isaac_b200.objectives import PandaReachObjective
    from oracle.backend import OracleBackend
    from scenes import panda_cfg
    single = MPPIisaacPlanner(panda_cfg(K=64, T=12), PandaReachObjective(), backend=OracleBackend())
    ref = np.stack([single.compute_action(Q0, [0] * 7).numpy() for _ in range(3)])
    a0, a1 = np.load(tmp_path / "act_0.npy"), np.load(tmp_path / "act_1.npy")
    np.testing.assert_array_equal(a0, a1)                               # every rank holds the same control
    np.testing.assert_allclose(a0, ref, atol=2e-6)                      # == unsharded plan (fp32 reassociation only)
    np.testing.assert_array_equal(np.load(tmp_path / "U_0.npy"), np.load(tmp_path / "U_1.npy"))
    sharded = np.concatenate([np.load(tmp_path / "actions_0.npy"), np.load(tmp_path / "actions_1.npy")], axis=2)
    np.testing.assert_allclose(sharded, single.mppi.actions.numpy(), atol=2e-6)   # global Philox index
