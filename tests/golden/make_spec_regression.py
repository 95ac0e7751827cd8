#!/usr/bin/env python
"""Regression fixture of the RESTATED SPEC (oracle/): spec_regression.json.

NOT a reference vector -- the reference's engines cannot run here (parity unpinned).  The fixture freezes what
the float64 oracle computes today for three small seeded cases (contact-free panda rollout, heijn push with contacts, the
softmax / update of both modes), so that a later change to the oracle or to the model compiler that alters the spec shows up as a
failing test instead of silently moving the target the CUDA kernels are compared with.   python tests/golden/make_spec_regression.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def cases():
    from oracle import oracle
    from scenes import panda_setup, push_setup
    out = {}
    sc, p, s0 = panda_setup(K=4, T=6)
    a = np.random.default_rng(11).uniform(-0.2, 0.2, (6, 7, 4)).astype(np.float32)
    st, obs = oracle.rollout(sc.model, p, s0, a, use_double=True)
    out["panda_rollout"] = dict(actions=a.tolist(), state=st.astype(np.float64).tolist(), obs_last=obs[:, -1].astype(np.float64).tolist())
    sc, p, s0 = push_setup(K=4, T=8, block_pos=(0.6, 1.5, 0.1))
    a = np.zeros((8, 3, 4), np.float32); a[:, 0] = np.array([0.6, 0.5, 0.4, 0.3], np.float32)
    st, obs = oracle.rollout(sc.model, p, s0, a, use_double=True, root0=sc.root_state0.astype(np.float32))
    out["heijn_push"] = dict(state=st.astype(np.float64).tolist(), obs_last=obs[:, -1].astype(np.float64).tolist())
    for mode in ("simple", "halton-spline"):
        sc, p, _ = panda_setup(K=8, T=12, mode=mode, filter_u=True)
        rng = np.random.default_rng(12)
        x = (rng.normal(size=(12, 7, 8)) * 0.1).astype(np.float32)
        cost = rng.uniform(0, 5, (12, 8)).astype(np.float32)
        U = (rng.normal(size=(12, 7)) * 0.05).astype(np.float32)
        part, _ = oracle.reduce(sc.model, p, cost, x, U)
        U2, act, stats = oracle.finalize(sc.model, p, part[None, :], U)
        out[f"update_{mode}"] = dict(x=x.tolist(), cost=cost.tolist(), U=U.tolist(), partial=np.asarray(part, np.float64).tolist(),
                                     U_new=np.asarray(U2, np.float64).tolist(), action=np.asarray(act, np.float64).tolist())
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "spec_regression.json"), "w") as fh:
        json.dump(dict(note="regression fixture of the restated spec (oracle/), not a reference vector", cases=cases()), fh)
    print("wrote spec_regression.json")
