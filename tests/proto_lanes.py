#!/usr/bin/env python
"""float64 numpy restatement of the substep of csrc/rollout_lanes.cu (frames by quaternion scan, composite rigid bodies,
joint-space LDL^T) checked against the oracle on the host -- a formulation check that needs no GPU.

    python tests/proto_lanes.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def qmul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def qrot(q, v):
    u = q[:3]
    c = 2 * np.cross(u, v)
    return v + q[3] * c + np.cross(u, c)


def q2R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def sym(v):
    return np.array([[v[0], v[3], v[4]], [v[3], v[1], v[5]], [v[4], v[5], v[2]]])


def rollout(m, p, state0, actions_k):
    """actions_k: (T, nu) of ONE rollout; returns (q, qd) after T steps."""
    nb, T = m.nb, p.T
    h = p.dt / p.substeps
    vel = m.drive_mode == 0
    q, qd = np.array(state0[:nb], float), np.array(state0[nb:2 * nb], float)
    bq = np.array(m.base_quat[:], float)
    bp = np.array(m.base_pos[:], float)
    a0f = -np.array(m.gravity[:], float) if m.gravity_on else np.zeros(3)
    tq = [np.array(m.tree_quat[i][:], float) for i in range(nb)]
    tp = [np.array(m.tree_p[i][:], float) for i in range(nb)]
    tR = [np.array(m.tree_R[i][:], float).reshape(3, 3) for i in range(nb)]
    rev = [m.jtype[i] == 0 for i in range(nb)]
    tax = [np.zeros(3) if rev[i] else tR[i][:, 2] for i in range(nb)]
    tp[0] = bp + qrot(bq, tp[0]); tax[0] = qrot(bq, tax[0]); tq[0] = qmul(bq, tq[0])
    mass = np.array(m.mass[:nb], float)
    com = [np.array(m.mcom[i][:], float) / mass[i] if mass[i] > 0 else np.zeros(3) for i in range(nb)]
    Ic = []
    for i in range(nb):
        Io = sym(np.array(m.inertia[i][:], float))
        c = com[i]
        Ic.append(Io - mass[i] * (c @ c * np.eye(3) - np.outer(c, c)))
    mc = np.cumsum(mass[::-1])[::-1]
    for t in range(T):
        u = actions_k[t] * p.u_scale
        tgt = np.array([m.cmd_c0[i] * u[m.cmd_i0[i]] + m.cmd_c1[i] * u[m.cmd_i1[i]] for i in range(nb)])
        for _ in range(p.substeps):
            # frames
            ql, pl = [], []
            for i in range(nb):
                ang = 0.5 * q[i] * (1.0 if rev[i] else 0.0)
                ql.append(qmul(tq[i], np.array([0, 0, np.sin(ang), np.cos(ang)])))
                pl.append(tp[i] + q[i] * tax[i])
            for i in range(1, nb):
                pl[i] = pl[i - 1] + qrot(ql[i - 1], pl[i]); ql[i] = qmul(ql[i - 1], ql[i])
            R = [q2R(x) for x in ql]
            Sn, Sf = [], []
            for i in range(nb):
                ax = R[i][:, 2]
                Sn.append(ax if rev[i] else np.zeros(3)); Sf.append(np.cross(pl[i], ax) if rev[i] else ax)
            Vn = np.cumsum([qd[i] * Sn[i] for i in range(nb)], axis=0)
            Vf = np.cumsum([qd[i] * Sf[i] for i in range(nb)], axis=0)
            A, hw, fn, ff_, an, af = [], [], [], [], [], []
            for i in range(nb):
                cw = pl[i] + R[i] @ com[i]
                hwi = mass[i] * cw
                Ai = R[i] @ Ic[i] @ R[i].T + (hwi @ cw) * np.eye(3) - np.outer(hwi, cw)
                w, v = Vn[i], Vf[i]
                nn = Ai @ w + np.cross(hwi, v)
                ff = mass[i] * v - np.cross(hwi, w)
                fn.append(np.cross(w, nn) + np.cross(v, ff)); ff_.append(np.cross(w, ff))
                an.append(np.cross(w, qd[i] * Sn[i])); af.append(np.cross(w, qd[i] * Sf[i]) + np.cross(v, qd[i] * Sn[i]))
                A.append(Ai); hw.append(hwi)
            an = np.cumsum(an, axis=0); af = np.cumsum(af, axis=0) + a0f
            for i in range(nb):
                fn[i] = fn[i] + A[i] @ an[i] + np.cross(hw[i], af[i])
                ff_[i] = ff_[i] + mass[i] * af[i] - np.cross(hw[i], an[i])
            Ac = np.cumsum(np.array(A)[::-1], axis=0)[::-1]
            hc = np.cumsum(np.array(hw)[::-1], axis=0)[::-1]
            fcn = np.cumsum(np.array(fn)[::-1], axis=0)[::-1]
            fcf = np.cumsum(np.array(ff_)[::-1], axis=0)[::-1]
            Fn = [Ac[i] @ Sn[i] + np.cross(hc[i], Sf[i]) for i in range(nb)]
            Ff = [mc[i] * Sf[i] - np.cross(hc[i], Sn[i]) for i in range(nb)]
            bias = np.array([Sn[i] @ fcn[i] + Sf[i] @ fcf[i] for i in range(nb)])
            M = np.zeros((nb, nb))
            for j in range(nb):
                for i in range(j + 1):
                    M[i, j] = M[j, i] = Sn[i] @ Fn[j] + Sf[i] @ Ff[j]
            sat = np.zeros(nb)
            for solve in range(2):
                tau, dimp = np.zeros(nb), np.zeros(nb)
                for i in range(nb):
                    kd, b, arm, eff = m.kd[i], m.damping[i], m.armature[i], m.effort[i]
                    if sat[i] != 0:
                        tau[i] = sat[i] * eff - b * qd[i]; dimp[i] = arm + h * b
                    elif vel:
                        tau[i] = kd * (tgt[i] - qd[i]) - b * qd[i]; dimp[i] = arm + h * (kd + b)
                    else:
                        tau[i] = min(max(tgt[i], -eff), eff) - (kd + b) * qd[i]; dimp[i] = arm + h * (kd + b)
                # distributed LDL^T as the kernel does it (column j of the upper triangle on lane j)
                col = [[M[r, j] + (dimp[j] if r == j else 0.0) for r in range(nb)] for j in range(nb)]   # col[j][r]
                lcol = [[0.0] * nb for _ in range(nb)]
                invd = np.ones(nb)
                for kk in range(nb):
                    dk = col[kk][kk]
                    inv = 1.0 / dk
                    lk = [col[j][kk] * inv for j in range(nb)]
                    invd[kk] = inv
                    for r in range(kk + 1, nb):
                        lr = lk[r]
                        for j in range(nb):
                            col[j][r] = col[j][r] - lr * col[j][kk]
                        lcol[kk][r] = lr
                    for j in range(kk + 1, nb):
                        col[j][kk] = lk[j]
                y = tau - bias
                for kk in range(nb - 1):
                    yk = y[kk]
                    for j in range(kk + 1, nb):
                        y[j] -= col[j][kk] * yk
                y = y * invd
                for jj in range(nb - 1, 0, -1):
                    xj = y[jj]
                    for i in range(jj):
                        y[i] -= lcol[i][jj] * xj
                qdd = y
                if solve == 0 and vel:
                    td = np.array([m.kd[i] * (tgt[i] - (qd[i] + h * qdd[i])) for i in range(nb)])
                    newly = np.abs(td) > np.array(m.effort[:nb])
                    if not newly.any():
                        break
                    sat = np.where(newly, np.sign(td), 0.0)
                else:
                    break
            for i in range(nb):
                vn = min(max(qd[i] + h * qdd[i], -m.qd_max[i]), m.qd_max[i])
                x = q[i] + h * vn
                if x < m.q_lo[i]:
                    x = m.q_lo[i]; vn = max(vn, 0.0)
                if x > m.q_hi[i]:
                    x = m.q_hi[i]; vn = min(vn, 0.0)
                q[i], qd[i] = x, vn
    return q, qd


def main():
    from oracle import oracle as orc
    from scenes import panda_setup, point_setup
    for name, setup, K, T in (("panda", panda_setup, 8, 30), ("point", point_setup, 8, 12)):
        sc, p, state0 = setup(K=K, T=T)
        rng = np.random.default_rng(1)
        lim = float(p.u_max[0])
        actions = rng.uniform(-lim, lim, (T, sc.nu, K)).astype(np.float32)
        if name == "panda":
            actions[:, :, 0] *= 8.0     # drives far beyond the effort limit -> saturation re-solve
        st_ref, _ = orc.rollout(sc.model, p, state0, actions, use_double=True)
        worst = 0.0
        for k in range(K):
            q, qd = rollout(sc.model, p, state0, actions[:, :, k].astype(float))
            nb = sc.model.nb
            worst = max(worst, np.abs(q - st_ref[:nb, k]).max(), np.abs(qd - st_ref[nb:2 * nb, k]).max() * 1e-2)
        print(f"{name}: max |q - oracle| (and 1e-2 |qd - oracle|) over {K} rollouts, T = {T}: {worst:.3e}")
        assert worst < 2e-5, name


if __name__ == "__main__":
    main()
