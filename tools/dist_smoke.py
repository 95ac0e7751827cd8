#!/usr/bin/env python
"""Step-by-step multi-GPU smoke (torchrun): NCCL init -> eager all_gather -> eager sharded plan -> graph-captured plan.
Prints a line after every stage (flush) so that a hang is attributable.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_smoke.py [eager|graph]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])


def say(msg):
    print(f"[rank {rank}] {time.strftime('%H:%M:%S')} {msg}", flush=True)


torch.cuda.set_device(local)
say("init_process_group")
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
say("eager all_gather")
x = torch.full((212,), float(rank), device=f"cuda:{local}")
out = torch.zeros((world, 212), device=f"cuda:{local}")
dist.all_gather_into_tensor(out.view(-1), x)
torch.cuda.synchronize()
assert out[:, 0].tolist() == [float(r) for r in range(world)]
say("all_gather ok")

import bench  # noqa: E402
from mppi_isaac_b200 import MPPIisaacPlanner  # noqa: E402
from mppi_isaac_b200.objectives import PandaReachObjective  # noqa: E402

graph = len(sys.argv) > 1 and sys.argv[1] == "graph"
planner = MPPIisaacPlanner(bench.panda_cfg(2000 * world, f"cuda:{local}"), PandaReachObjective(), use_cuda_graph=graph)
say(f"planner built: K_local={planner.sim.num_envs} k_offset={planner.k_offset} graph={graph} peer_exchange={planner.mppi._peer_exchange}")
q0, goal = bench.synthetic_state()
planner.sim.set_actor_position_by_name(goal, "goal")
planner.sim.reset_robot_state(q0, np.zeros(7))
for i in range(4):
    a = planner.mppi.command()
    torch.cuda.synchronize()
    say(f"plan {i}: action[0]={float(a[0]):+.5f} graph_captured={planner.mppi._graph is not None}")
acts = [torch.zeros_like(a) for _ in range(world)]
dist.all_gather(acts, a.contiguous())
assert all(torch.equal(acts[0], t) for t in acts), "ranks disagree on the action"
say("all ranks hold the same action")
if rank == 0:
    print("ACTION " + ("peer" if planner.mppi._peer_exchange else "nccl") + " " + " ".join(f"{float(v):.9e}" for v in a), flush=True)
bench.shutdown_distributed(planner)
say("done")
