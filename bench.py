#!/usr/bin/env python
"""bench.py -- plan-loop Hz / rollout-steps per second of the MPPI rollout hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE MPPI plan of the BASELINE headline workload (panda 7-DoF reach, K = 10 000 samples per GPU,
T = 30, dt 0.05 / 2 substeps): shift U -> K1 sample/clamp -> K2 articulated rollout -> Objective cost ->
K3 fused cost/softmax/weighted sum -> [NCCL all-gather of the shard partials] -> K4 update.
Samples are sharded over the ranks (weak scaling: every GPU owns K = 10 000 samples, K_total = N * 10 000).

Printed JSON (rank 0, one line):
  value      rollout-steps/s (= K_total * T * plans/s), inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        the same metric through MPPIisaacPlanner.compute_action_tensor(dof_bytes, root_bytes) with host buffers
  roofline   K3 (fused cost-softmax-weighted-sum) achieved HBM GB/s vs the measured peak (MEASURED_PEAKS.json)
  cpu_baseline  the CPU restatement of the reference pipeline (oracle/) timed on this box's host cores (N=1 only)
--impl reference times that CPU restatement as the reference arm (the reference's own engines, IsaacGym/PhysX and
mppi_torch, are closed / un-vendored and cannot run here: BASELINE.md section 2).
"""
import argparse
import copy
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

K_PER_GPU = 10000
T_HORIZON = 30
WORKLOAD = "panda 7-DoF reach (BASELINE C2*): K=10000/GPU, T=30, dt=0.05, substeps=2, Gaussian sampling, cost O1 (PandaReachObjective, fused ops.pose_cost)"
METRIC = "rollout_steps_per_sec"
UNIT = "rollout-steps/s"


def panda_cfg(K, device):
    from mppi_isaac_b200.utils.config_store import load_isaacgym_config
    cfg = copy.deepcopy(load_isaacgym_config("config_panda_b200"))
    cfg.mppi.num_samples, cfg.mppi.horizon, cfg.mppi.device = int(K), T_HORIZON, device
    return cfg


def synthetic_state(seed=1234 + 2):
    """q0 ~ U(lower+0.1, upper-0.1), qd0 = 0, goal ~ U([0.3,0.7]x[-0.4,0.4]x[0.2,0.7])."""
    g = np.random.default_rng(seed)
    lo = np.array([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973]) + 0.1
    hi = np.array([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973]) - 0.1
    q0 = g.uniform(lo, hi)
    goal = g.uniform([0.3, -0.4, 0.2], [0.7, 0.4, 0.7])
    return q0, goal


def world_bytes(planner, q, qd, goal):
    from mppi_isaac_b200.utils.transport import torch_to_bytes
    dof = torch.tensor([[v for a, b in zip(q, qd) for v in (a, b)]], dtype=torch.float32)
    root = torch.from_numpy(planner.sim.scene.root_state0.copy()).unsqueeze(0)
    root[0, planner.sim._get_actor_index_by_name("goal"), 0:3] = torch.tensor(goal, dtype=torch.float32)
    return torch_to_bytes(dof), torch_to_bytes(root), dof.numel() * 4 + root.numel() * 4


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------------
# CPU restatement arm (cpu_baseline and --impl reference)
# --------------------------------------------------------------------------------------------------
def cpu_plan_rate(k_sample, steps, warmup, cores):
    """Plans of the oracle pipeline (sample -> rollout -> Objective -> reduce -> finalize) on `cores` host threads."""
    from mppi_isaac_b200 import MPPIisaacPlanner
    from mppi_isaac_b200.objectives import PandaReachObjective
    from oracle.backend import OracleBackend
    torch.set_num_threads(max(1, min(8, cores)))            # the Objective's torch ops; more threads only add overhead here
    planner = MPPIisaacPlanner(panda_cfg(k_sample, "cpu"), PandaReachObjective(), backend=OracleBackend(nthreads=min(cores, 64)))
    q0, goal = synthetic_state()
    planner.sim.set_actor_position_by_name(goal, "goal")
    planner.sim.reset_robot_state(q0, np.zeros(7))
    for _ in range(warmup):
        planner.mppi.command()
    t0 = time.perf_counter()
    for _ in range(steps):
        planner.mppi.command()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dt


def pick_cpu_sample(cores, budget_s):
    """Largest K (<= 10 000, multiple of 4) whose single plan fits `budget_s` on this host, from a K=256 probe."""
    probe = cpu_plan_rate(512, 2, 1, cores)
    per_sample = probe / 512
    k = int(min(K_PER_GPU, max(256, budget_s / per_sample)))
    return max(256, (k // 4) * 4)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    import __graft_entry__
    from oracle import oracle as orc
    orc.build()
    cores = os.cpu_count() or 1
    k_s = pick_cpu_sample(cores, budget_s=2.0)
    dt = cpu_plan_rate(k_s, args.steps, args.warmup, cores)
    value = k_s * T_HORIZON / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "plan_hz_at_sample": 1.0 / dt, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU restatement of the reference pipeline (oracle/), not IsaacGym/PhysX: those cannot run here"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} plans of K={k_s} of the K=10000 workload (rate is per rollout-step, K-independent)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def graph_time_us(fn, reps, replays=3):
    """Device time of `fn` (us per call): `reps` back-to-back calls captured in ONE CUDA graph, so that the
    measurement is free of Python/ctypes launch overhead; CUDA events on the replaying stream, best of `replays`."""
    fn(); torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    best = float("inf")
    for _ in range(replays):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e3 / reps)
    return best


def time_kernels(planner, reps=20):
    """Per-kernel device time (us) of one plan (graph-captured back-to-back launches, warm L2)."""
    m, be, sim = planner.mppi, planner.mppi.backend, planner.sim
    from mppi_isaac_b200.model.blob import MODE_SIMPLE
    x = m.noise if be.params.mode == MODE_SIMPLE else m.actions

    def t(fn):
        return graph_time_us(fn, reps)

    out = {
        "sample_us": t(lambda: be.sample(m.seed, 0, m.k_offset, m.K_total, m.U, None, m.actions, m.noise, m.plan_ctr)),
        "rollout_us": t(lambda: sim.rollout_all(m.actions)),
        "cost_objective_us": t(lambda: m._cost_batched()),      # the Objective as benchmarked (one fused ops.pose_cost launch)
    }
    obj = planner.objective
    if getattr(obj, "fused", False):
        obj.fused = False
        out["cost_objective_torch_ops_us"] = t(lambda: m._cost_batched())   # same Objective written with ~28 torch launches
        obj.fused = True
    cost = m._cost_batched()
    out["reduce_us_warm_l2"] = t(lambda: be.reduce(cost, x, m.U, m.partial))
    u_tmp = m.U.clone()
    out["finalize_us"] = t(lambda: be.finalize(m.partial.view(1, -1), 1, u_tmp, m._action, m.stats))
    if m.world == 1:        # what a single-GPU plan actually launches: K3 with K4 done by its last CTA
        out["reduce_finalize_fused_us_warm_l2"] = t(lambda: be.reduce_finalize(cost, x, u_tmp, m.partial, m._action, m.stats))
    return out


def k3_roofline(planner, peak_gbs, peak_src, K_list):
    """K3 alone, inputs rotated over > L2 worth of distinct buffers so every launch reads HBM (cold L2)."""
    from mppi_isaac_b200.backend import CudaBackend
    from mppi_isaac_b200.model.blob import MppibParams
    dev = planner.sim.device
    T, nu = planner.mppi.T, planner.mppi.nu
    res = []
    for K in K_list:
        p = MppibParams.from_buffer_copy(bytes(planner.mppi.backend.params))
        p.K = K
        be = CudaBackend(dev)
        be.create(planner.sim.scene.model, p)
        bytes_alg = 4 * K * T * (nu + 1) + 4 * (T * nu + 2)
        nbuf = max(2, int(np.ceil(300e6 / bytes_alg)))            # > 2x the 126 MB L2
        nbuf = min(nbuf, 64)
        xs = [torch.randn((T, nu, K), device=dev) * 0.3 for _ in range(nbuf)]
        cs = [torch.rand((T, K), device=dev) * 10 for _ in range(nbuf)]
        U = torch.zeros((T, nu), device=dev)
        partial = torch.zeros(2 + T * nu, device=dev)
        def sweep():
            for i in range(nbuf):
                be.reduce(cs[i], xs[i], U, partial)
        us = graph_time_us(sweep, 2) / nbuf
        gbs = bytes_alg / (us * 1e-6) / 1e9
        res.append({"K": K, "bytes": bytes_alg, "us": us, "GBps": gbs, "frac": gbs / peak_gbs, "l2": f"cold: {nbuf} rotating input sets ({nbuf * bytes_alg / 1e6:.0f} MB)"})
        del xs, cs
        be.destroy()
    return res


def run_gpu_arm(args, rank, world, local_rank):
    import torch.distributed as dist
    import __graft_entry__
    __graft_entry__.build()
    from mppi_isaac_b200 import MPPIisaacPlanner
    from mppi_isaac_b200.objectives import PandaReachObjective
    from mppi_isaac_b200.utils.transport import bytes_to_torch

    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    k_total = K_PER_GPU * world
    planner = MPPIisaacPlanner(panda_cfg(k_total, dev), PandaReachObjective(), use_cuda_graph=True)
    assert planner.sim.num_envs == K_PER_GPU
    q0, goal = synthetic_state()
    dof_b, root_b, h2d = world_bytes(planner, q0, np.zeros(7), goal)
    planner.objective.reset()
    planner.reset_rollout_sim(dof_b, root_b)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2
    # ---- device-resident timing: K plans, CUDA events around each plan, L2 flushed between plans -------------------
    for _ in range(max(args.warmup, 3)):
        planner.mppi.command()
    graph_on = planner.mppi._graph is not None
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # before the barrier: spawning nvidia-smi takes milliseconds, and with the exchange fused
    barrier()                    # into the kernels every rank's first timed plan would wait for a late rank 0
    for i in range(args.steps):
        flush.zero_()
        starts[i].record()
        planner.mppi.command()
        ends[i].record()
    barrier()
    per_step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = torch.tensor([sum(per_step_ms)], dtype=torch.float64, device=dev)
    # ---- end to end through the plugin API with host buffers ----------------------------------------------------
    # the caller's side of the wire (building / pickling the world state) is prepared outside the timed region: a
    # different synthetic joint state per step, in the reference's own torch.save byte format (transport.py:5-14)
    rng = np.random.default_rng(99)
    inputs = [world_bytes(planner, q0 + rng.uniform(-0.05, 0.05, 7), rng.uniform(-0.1, 0.1, 7), goal)[:2] for _ in range(args.steps)]
    for _ in range(3):
        bytes_to_torch(planner.compute_action_tensor(dof_b, root_b))
    barrier()
    outs = []
    t0 = time.perf_counter()
    for i in range(args.steps):
        outs.append(planner.compute_action_tensor(*inputs[i]))     # bytes in -> H2D -> plan -> D2H -> bytes out
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    assert all(bytes_to_torch(o).shape == (planner.mppi.nu,) for o in outs)
    # the reference's in-process entry point: compute_action(q, qdot) with host lists in, host tensor out
    qs = [list(q0 + rng.uniform(-0.05, 0.05, 7)) for _ in range(args.steps)]
    qds = [list(rng.uniform(-0.1, 0.1, 7)) for _ in range(args.steps)]
    for _ in range(3):
        planner.compute_action(qs[0], qds[0])
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        planner.compute_action(qs[i], qds[i])
    torch.cuda.synchronize()
    e2e2_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e2_s, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) * 1e-3
    value = k_total * T_HORIZON * args.steps / total_s
    e2e_value = k_total * T_HORIZON * args.steps / float(e2e_s.item())

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        launches_per_plan = (5 if world == 1 else 6)   # shift, sample, rollout, fused pose cost (Objective), reduce(+finalize fused at 1 GPU) [, finalize]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_s * 1e3 / args.steps, "plan_hz": args.steps / total_s, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "K_per_gpu": K_PER_GPU, "K_total": k_total, "T": T_HORIZON, "parallelism": f"sample-shard x{world}",
                       "exchange": "none" if world == 1 else ("peer-memory stores fused into K3 (NVLink), flags acquired by K4" if planner.mppi._peer_exchange else "NCCL all-gather"),
                       "cuda_graph": graph_on, "l2": "flushed (256 MiB write) before every timed plan",
                       "ms_per_step_p10_p50_p90_max": [float(np.percentile(per_step_ms, p)) for p in (10, 50, 90, 100)]},
            "e2e": {"value": e2e_value, "unit": UNIT, "plan_hz": args.steps / float(e2e_s.item()), "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": planner.mppi.nu * 4, "api": "MPPIisaacPlanner.compute_action_tensor(dof_bytes, root_bytes) -> bytes",
                    "compute_action_plan_hz": args.steps / float(e2e2_s.item())},
            "gpu_launches": launches_per_plan * args.steps,
            "clocks": clocks,
        }
        if world == 1:
            # transparency: the same plan with the Objective written as plain torch ops (~28 element-wise launches instead of the
            # one fused ops.pose_cost launch) -- what an unmodified user Objective costs
            planner.objective.fused = False
            planner.mppi.invalidate_graph()
            for _ in range(3):
                planner.mppi.command()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(args.steps, 30))]
            for a, b in ev:
                flush.zero_()
                a.record()
                planner.mppi.command()
                b.record()
            torch.cuda.synchronize()
            ms_torch_obj = float(np.mean([a.elapsed_time(b) for a, b in ev]))
            line["objective_as_torch_ops"] = {"ms_per_step": ms_torch_obj, "plan_hz": 1e3 / ms_torch_obj,
                                              "value": k_total * T_HORIZON / (ms_torch_obj * 1e-3), "unit": UNIT}
            planner.objective.fused = True
            planner.mppi.invalidate_graph()
            planner.mppi.command()
            kt = time_kernels(planner)
            roof = k3_roofline(planner, peak, peak_src, [K_PER_GPU, 65536, 262144])
            head = roof[0]
            line["roofline"] = {"kernel": "K3 reduce_kernel (fused cost accumulate + softmax + weighted control sum)", "bound": "hbm",
                                "achieved": head["GBps"], "peak": peak, "unit": "GB/s", "frac": head["frac"],
                                "traffic": 9635000 if (head["K"], T_HORIZON) == (10000, 30) else None,
                                "traffic_source": "ncu --set full dram__bytes_read.sum + write of one K3 launch at K=10000 (profiles/r1_reduce_v2.md): 1.004 x algorithmic",
                                "peak_source": peak_src, "bytes_per_launch": head["bytes"], "us_per_launch": head["us"], "K": head["K"],
                                "note": "9.6 MB per launch at the named K is ~1.5 us of HBM time, i.e. launch/latency bound; see sweep for the asymptote",
                                "sweep": roof}
            line["kernels_us"] = kt
            cores = os.cpu_count() or 1
            k_s = pick_cpu_sample(cores, budget_s=2.0)
            n_cpu = 5
            dt = cpu_plan_rate(k_s, n_cpu, 1, cores)
            line["cpu_baseline"] = {"value": k_s * T_HORIZON / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{n_cpu} plans of K={k_s} of the K=10000 workload, CPU restatement (oracle/) on {cores} threads",
                                    "plan_hz_at_K10000_est": 1.0 / (dt * K_PER_GPU / k_s)}
            # one host thread: how the reference configures PhysX (num_threads never set, isaacgym_wrapper.py:21-39)
            dt1 = cpu_plan_rate(256, 3, 1, 1)
            line["cpu_baseline"]["one_thread"] = {"value": 256 * T_HORIZON / dt1, "unit": UNIT, "sample": "3 plans of K=256 on 1 thread",
                                                  "plan_hz_at_K10000_est": 1.0 / (dt1 * K_PER_GPU / 256)}
        print(json.dumps(line), flush=True)
    if world > 1:
        shutdown_distributed(planner)


def shutdown_distributed(planner):
    """Tear NCCL down without hanging: a live CUDA graph that holds NCCL kernels blocks destroy_process_group(), so the
    captured plan is released first; a watchdog force-exits if the teardown still does not return."""
    import gc
    import torch.distributed as dist
    planner.mppi.invalidate_graph()
    planner.mppi.close_peers()          # collective: unmap the peer-memory exchange windows before the group goes away
    del planner
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    timer = threading.Timer(20.0, lambda: os._exit(0))
    timer.daemon = True
    timer.start()
    dist.destroy_process_group()
    timer.cancel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference for the CPU arm)")
    run_gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
