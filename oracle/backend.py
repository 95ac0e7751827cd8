"""Checker backend: the CPU oracle behind the same interface as ``mppi_isaac_b200.backend.CudaBackend``.

TEST INFRASTRUCTURE (also the CPU-baseline arm of bench.py).  It lets the `-m "not gpu"` suite drive the host logic (planner sequencing, facade views,
sharding, transport) end to end on CPU tensors, and gives the GPU parity tests a planner-level reference.  The
product package never imports this file or ``oracle/``.
"""
import numpy as np
import torch

from . import oracle as orc


def _np(t):
    return None if t is None else t.detach().numpy()


class OracleBackend:
    name = "oracle"

    def __init__(self, device="cpu", use_double=False, nthreads=1):
        self.device = torch.device("cpu")
        self.model = self.params = None
        self.use_double, self.nthreads = use_double, nthreads
        self.launches = 0

    def create(self, model, params):
        self.model, self.params = model, params

    def destroy(self):
        pass

    def set_params(self, params):
        self.params = params

    def set_model(self, model):
        self.model = model

    def state_size(self):
        return 2 * self.model.nb + 13 * self.model.nfree

    def obs_size(self):
        return orc.obs_size(self.model, self.params)

    def sample(self, seed, plan_idx, k_offset, k_total, U, prior_row, actions, noise, plan_ctr=None):
        plan = plan_idx + (int(plan_ctr[0]) if plan_ctr is not None else 0)
        a, n = orc.sample(self.model, self.params, seed, plan, _np(U), k_offset, k_total, _np(prior_row), self.nthreads)
        actions.copy_(torch.from_numpy(a))
        if noise is not None:
            noise.copy_(torch.from_numpy(n))

    def noise_library(self, k_offset, k_total, halton_tab, B, n_knots, Z):
        Z.copy_(torch.from_numpy(orc.noise_library(self.model, self.params, _np(halton_tab), _np(B), n_knots, k_offset, k_total)))

    def sample_library(self, k_offset, k_total, U, prior_row, Z, actions, noise):
        a, n = orc.sample_library(self.model, self.params, _np(U), _np(Z), k_offset, k_total, _np(prior_row))
        actions.copy_(torch.from_numpy(a))
        if noise is not None:
            noise.copy_(torch.from_numpy(n))

    def rollout(self, state0, state, actions, t0, nsteps, obs, act_t0=0, root0=None):
        T, K, nu = self.params.T, self.params.K, self.model.nu
        src = _np(actions).reshape(-1, nu, K)
        if act_t0 == 0 and src.shape[0] == T and src.flags["C_CONTIGUOUS"] and src.dtype == np.float32:
            a = src                                     # whole-horizon launch: the caller's buffer as it is
        else:
            a = np.zeros((T, nu, K), np.float32)
            a[act_t0:act_t0 + src.shape[0]] = src
        st = _np(state) if state is not None else None
        s0 = _np(state0)
        import ctypes as C
        f = lambda x: None if x is None else x.ctypes.data_as(C.POINTER(C.c_float))
        o = _np(obs)
        assert o is None or o.flags["C_CONTIGUOUS"]
        assert st is None or st.flags["C_CONTIGUOUS"]
        r0 = None if root0 is None else np.ascontiguousarray(_np(root0), np.float32)
        orc.lib().oracle_rollout(C.byref(self.model), C.byref(self.params), f(None if s0 is None else np.ascontiguousarray(s0, np.float32)),
                                 f(r0), f(st), f(a), C.c_int32(t0), C.c_int32(nsteps), f(o), C.c_int32(int(self.use_double)), C.c_int32(self.nthreads))

    def reduce(self, cost, x, U, partial):
        if self.nthreads > 1:          # timing arm: threaded, no staging copies (the single-threaded form below is the parity checker)
            c = cost if cost.is_contiguous() else cost.contiguous()
            partial.copy_(torch.from_numpy(orc.reduce_mt(self.model, self.params, _np(c), _np(x), _np(U), self.nthreads)))
            return
        p, _ = orc.reduce(self.model, self.params, _np(cost.contiguous()), _np(x), _np(U))
        partial.copy_(torch.from_numpy(p))

    def finalize(self, partials, G, U, action_out, stats):
        Un, act, st = orc.finalize(self.model, self.params, _np(partials.contiguous())[:G], _np(U))
        U.copy_(torch.from_numpy(Un))
        action_out.copy_(torch.from_numpy(act))
        if stats is not None:
            stats.copy_(torch.from_numpy(st))

    def shift(self, U, plan_ctr=None):
        U.copy_(torch.from_numpy(orc.shift(self.model, self.params, _np(U))))
        if plan_ctr is not None:
            plan_ctr += 1


class OraclePandaReachObjective:
    """The panda reach Objective (O1) of the CPU timing arm: the same two terms as ``PandaReachObjective`` evaluated by the
    oracle's threaded ``oracle_cost_pose`` -- the CPU counterpart of the fused ``ops.pose_cost`` kernel the GPU arm uses."""

    def __init__(self, nthreads=1, actor="panda", link="panda_ee_tip", goal="goal"):
        self.weights = {"robot_to_goal": 1.0, "robot_ori": 0.5}
        self.actor, self.link, self.goal, self.nthreads = actor, link, goal, nthreads
        self._out = None

    def reset(self):
        pass

    def compute_cost(self, sim):
        ee = sim.get_actor_link_by_name(self.actor, self.link)
        goal = sim.get_actor_position_by_name(self.goal)
        n = ee.shape[0]
        if self._out is None or self._out.shape[0] != n:
            self._out = torch.empty((n,), dtype=torch.float32)
        orc.cost_pose(_np(ee), _np(goal), self.weights["robot_to_goal"], self.weights["robot_ori"], _np(self._out), self.nthreads)
        return self._out
