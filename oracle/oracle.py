"""ctypes front-end of ``liboracle.so`` (numpy in / numpy out).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import this module.  PARITY UNPINNED (see oracle.cpp header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from mppi_isaac_b200.model.blob import MppibModel, MppibParams

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "mppib.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_obs_size.restype = C.c_int32
        _LIB.oracle_state_size.restype = C.c_int32
        _LIB.oracle_halton.restype = C.c_double
        _LIB.oracle_norm_ppf.restype = C.c_double
    return _LIB


def _f(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def philox(counter, key):
    out = (C.c_uint32 * 4)()
    lib().oracle_philox(*[C.c_uint32(int(c)) for c in counter], C.c_uint32(int(key[0])), C.c_uint32(int(key[1])), out)
    return list(out)


def obs_size(model: MppibModel, params: MppibParams) -> int:
    return lib().oracle_obs_size(C.byref(model), C.byref(params))


def sample(model, params, seed, plan_idx, U, k_offset=0, k_total=None, prior_row=None, nthreads=1):
    K, T, nu = params.K, params.T, model.nu
    k_total = K if k_total is None else k_total
    U = np.ascontiguousarray(U, np.float32).reshape(T, nu)
    actions = np.empty((T, nu, K), np.float32)
    noise = np.empty((T, nu, K), np.float32)
    pr = None if prior_row is None else np.ascontiguousarray(prior_row, np.float32)
    lib().oracle_sample(C.byref(model), C.byref(params), C.c_uint64(seed), C.c_uint64(plan_idx), C.c_uint32(k_offset),
                        C.c_uint32(k_total), _f(U), _f(pr), _f(actions), _f(noise), C.c_int(nthreads))
    return actions, noise


def noise_library(model, params, halton_tab, B, n_knots, k_offset=0, k_total=None):
    K, T, nu = params.K, params.T, model.nu
    k_total = K if k_total is None else k_total
    tab = np.ascontiguousarray(halton_tab, np.int32)
    Bm = np.ascontiguousarray(B, np.float32)
    Z = np.zeros((T, nu, K), np.float32)
    lib().oracle_noise_library(C.byref(model), C.byref(params), C.c_uint32(k_offset), C.c_uint32(k_total), tab.ctypes.data_as(C.POINTER(C.c_int32)),
                               _f(Bm), C.c_int32(n_knots), _f(Z))
    return Z


def sample_library(model, params, U, Z, k_offset=0, k_total=None, prior_row=None):
    K, T, nu = params.K, params.T, model.nu
    k_total = K if k_total is None else k_total
    U = np.ascontiguousarray(U, np.float32).reshape(T, nu)
    Z = np.ascontiguousarray(Z, np.float32)
    actions, noise = np.empty((T, nu, K), np.float32), np.empty((T, nu, K), np.float32)
    pr = None if prior_row is None else np.ascontiguousarray(prior_row, np.float32)
    lib().oracle_sample_library(C.byref(model), C.byref(params), C.c_uint32(k_offset), C.c_uint32(k_total), _f(U), _f(pr), _f(Z), _f(actions), _f(noise))
    return actions, noise


def rollout(model, params, state0, actions, t0=0, nsteps=None, state=None, want_obs=True, use_double=False, nthreads=1, root0=None):
    """state0: (2*ndof,) broadcast row or None (continue from `state` (NS,K) in place); root0: (A,13) actor root states."""
    K, T = params.K, params.T
    nsteps = T if nsteps is None else nsteps
    NS = lib().oracle_state_size(C.byref(model))
    R = obs_size(model, params)
    obs = np.zeros((R, T, K), np.float32) if want_obs else None
    if state is None:
        state = np.zeros((NS, K), np.float32)
    s0 = None if state0 is None else np.ascontiguousarray(state0, np.float32)
    actions = np.ascontiguousarray(actions, np.float32)
    r0 = None if root0 is None else np.ascontiguousarray(root0, np.float32)
    lib().oracle_rollout(C.byref(model), C.byref(params), _f(s0), _f(r0), _f(state), _f(actions), C.c_int32(t0), C.c_int32(nsteps),
                         _f(obs), C.c_int32(int(use_double)), C.c_int32(nthreads))
    return state, obs


def reduce(model, params, cost, x, U):
    K, T, nu = params.K, params.T, model.nu
    partial = np.zeros(2 + T * nu, np.float32)
    S = np.zeros(K, np.float32)
    lib().oracle_reduce(C.byref(model), C.byref(params), _f(np.ascontiguousarray(cost, np.float32)),
                        _f(np.ascontiguousarray(x, np.float32)), _f(np.ascontiguousarray(U, np.float32)), _f(partial), _f(S))
    return partial, S


def finalize(model, params, partials, U):
    T, nu = params.T, model.nu
    partials = np.ascontiguousarray(partials, np.float32).reshape(-1, 2 + T * nu)
    U = np.array(U, np.float32).reshape(T, nu).copy()
    action = np.zeros(nu, np.float32)
    stats = np.zeros(2, np.float32)
    lib().oracle_finalize(C.byref(model), C.byref(params), _f(partials), C.c_int32(partials.shape[0]), _f(U), _f(action), _f(stats))
    return U, action, stats


def shift(model, params, U):
    U = np.array(U, np.float32).reshape(params.T, model.nu).copy()
    lib().oracle_shift(C.byref(model), C.byref(params), _f(U))
    return U


def reduce_mt(model, params, cost, x, U, nthreads):
    """Threaded variant for the CPU timing arm (same rule, per-thread partials merged); numpy views, no copies."""
    T, nu = params.T, model.nu
    partial = np.zeros(2 + T * nu, np.float32)
    lib().oracle_reduce_mt(C.byref(model), C.byref(params), _f(cost), _f(x), _f(U), _f(partial), C.c_int32(int(nthreads)))
    return partial


def cost_pose(a, b, w_pos, w_ori, out, nthreads):
    """a: (N, >=7) float32 strided view, b: (N, >=3) strided view or None; out: contiguous (N,) float32 (numpy arrays)."""
    n = a.shape[0]
    ap = a.ctypes.data_as(C.POINTER(C.c_float))
    if b is None:
        bp, bsi, bsr = None, 0, 0
    else:
        bp, bsi, bsr = b.ctypes.data_as(C.POINTER(C.c_float)), b.strides[0] // 4, b.strides[1] // 4
    lib().oracle_cost_pose(C.c_int64(n), ap, C.c_int64(a.strides[0] // 4), C.c_int64(a.strides[1] // 4), bp, C.c_int64(bsi), C.c_int64(bsr),
                           C.c_float(w_pos), C.c_float(w_ori), _f(out), C.c_int32(int(nthreads)))
    return out
