#!/usr/bin/env python
"""float64 numpy restatement of the ARTICULATION substep of csrc/rollout_team.cu for TREES, written the way the kernel computes it --
world frames / velocities / accelerations by pointer jumping over the ancestors, composites as differences of suffix sums over the
depth-first body order, joint-space LDL^T with the leaves eliminated first -- and checked against the oracle (body-frame ABA with
dense 6x6 transforms) on the host; and of its contact solve over generalised coordinates (a free box on the ground: linear + BODY-axis
angular velocity components with scalar inverse inertias, cached rows) against the oracle's world-frame solve.  Formulation checks
that need no GPU (tests/test_proto_team.py runs them):

    python tools/proto_team.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from proto_lanes import q2R, qmul, qrot, sym  # noqa: E402


def tree_tables(parent, G):
    """What the kernel keeps per lane: jump[r][i] = the 2^r-th ancestor of body i (-1: none), desc[i] = set of i and its descendants,
    end[i] = first body after the subtree of i (depth-first numbering: the subtree is the index range [i, end))."""
    nb = len(parent)
    R = {4: 2, 8: 3, 16: 4}[G]
    jump = [list(parent)]
    for r in range(1, R):
        jump.append([jump[r - 1][jump[r - 1][i]] if jump[r - 1][i] >= 0 else -1 for i in range(nb)])
    anc = [{i} for i in range(nb)]
    for r in range(R):
        anc = [anc[i] | (anc[jump[r][i]] if jump[r][i] >= 0 else set()) for i in range(nb)]
    desc = [{j for j in range(nb) if i in anc[j]} for i in range(nb)]
    end = [i + len(desc[i]) for i in range(nb)]
    for i in range(nb):
        assert desc[i] == set(range(i, end[i])), "bodies must be numbered depth first"
        assert len(anc[i]) <= 2 ** R, "tree deeper than the team is wide"
    return jump, desc, end


def anc_sum(x, jump):
    """x[i] <- sum of x over i and its ancestors (pointer jumping: after round r every entry covers 2^(r+1) bodies up the tree)"""
    x = [np.array(v, float) for v in x]
    for jr in jump:
        x = [x[i] + x[jr[i]] if jr[i] >= 0 else x[i] for i in range(len(x))]
    return x


def subtree_sum(x, end):
    """x[i] <- sum over the subtree of i = suffix sum at i minus suffix sum at the subtree's end"""
    P = np.cumsum(np.array(x, float)[::-1], axis=0)[::-1]
    return [P[i] - (P[end[i]] if end[i] < len(x) else 0.0) for i in range(len(x))]


def rollout(m, p, state0, actions_k, want_pivots=False):
    """actions_k: (T, nu) of ONE rollout; returns (q, qd) after T steps [and the LDL pivots of the last substep]."""
    nb, T = m.nb, p.T
    G = 8 if nb <= 8 else 16
    h = p.dt / p.substeps
    vel = m.drive_mode == 0
    parent = [m.parent[i] for i in range(nb)]
    jump, desc, end = tree_tables(parent, G)
    q, qd = np.array(state0[:nb], float), np.array(state0[nb:2 * nb], float)
    bq, bp = np.array(m.base_quat[:], float), np.array(m.base_pos[:], float)
    a0f = -np.array(m.gravity[:], float) if m.gravity_on else np.zeros(3)
    tq = [np.array(m.tree_quat[i][:], float) for i in range(nb)]
    tp = [np.array(m.tree_p[i][:], float) for i in range(nb)]
    tR = [np.array(m.tree_R[i][:], float).reshape(3, 3) for i in range(nb)]
    rev = [m.jtype[i] == 0 for i in range(nb)]
    tax = [np.zeros(3) if rev[i] else tR[i][:, 2] for i in range(nb)]
    for i in range(nb):
        if parent[i] < 0:       # the base pose is folded into every root body's parent transform
            tp[i] = bp + qrot(bq, tp[i]); tax[i] = qrot(bq, tax[i]); tq[i] = qmul(bq, tq[i])
    mass = np.array(m.mass[:nb], float)
    com = [np.array(m.mcom[i][:], float) / mass[i] if mass[i] > 0 else np.zeros(3) for i in range(nb)]
    Ic = []
    for i in range(nb):
        c = com[i]
        Ic.append(sym(np.array(m.inertia[i][:], float)) - mass[i] * (c @ c * np.eye(3) - np.outer(c, c)))
    mc = subtree_sum(list(mass), end)
    pivots = None
    for t in range(T):
        u = actions_k[t] * p.u_scale
        tgt0 = np.array([m.cmd_c0[i] * u[m.cmd_i0[i]] + m.cmd_c1[i] * u[m.cmd_i1[i]] for i in range(nb)])
        for _ in range(p.substeps):
            tgt = tgt0.copy()
            if m.planar_base:   # body twist (v, omega) -> world-frame velocity targets of the three virtual joints
                sy, cy = np.sin(q[2]), np.cos(q[2])
                tgt[0] = u[0] * (m.fwd_axis[0] * cy - m.fwd_axis[1] * sy)
                tgt[1] = u[0] * (m.fwd_axis[0] * sy + m.fwd_axis[1] * cy)
                tgt[2] = u[1]
            # frames: product of the local transforms along the path from the root, by pointer jumping
            ql, pl = [], []
            for i in range(nb):
                ang = 0.5 * q[i] * (1.0 if rev[i] else 0.0)
                ql.append(qmul(tq[i], np.array([0, 0, np.sin(ang), np.cos(ang)])))
                pl.append(tp[i] + q[i] * tax[i])
            for jr in jump:
                nq, npos = list(ql), list(pl)
                for i in range(nb):
                    if jr[i] >= 0:
                        npos[i] = pl[jr[i]] + qrot(ql[jr[i]], pl[i]); nq[i] = qmul(ql[jr[i]], ql[i])
                ql, pl = nq, npos
            R = [q2R(x) for x in ql]
            Sn, Sf = [], []
            for i in range(nb):
                ax = R[i][:, 2]
                Sn.append(ax if rev[i] else np.zeros(3)); Sf.append(np.cross(pl[i], ax) if rev[i] else ax)
            Vn = anc_sum([qd[i] * Sn[i] for i in range(nb)], jump)
            Vf = anc_sum([qd[i] * Sf[i] for i in range(nb)], jump)
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
            an = anc_sum(an, jump); af = [x + a0f for x in anc_sum(af, jump)]
            for i in range(nb):
                fn[i] = fn[i] + A[i] @ an[i] + np.cross(hw[i], af[i])
                ff_[i] = ff_[i] + mass[i] * af[i] - np.cross(hw[i], an[i])
            Ac, hc, fcn, fcf = subtree_sum(A, end), subtree_sum(hw, end), subtree_sum(fn, end), subtree_sum(ff_, end)
            Fn = [Ac[i] @ Sn[i] + np.cross(hc[i], Sf[i]) for i in range(nb)]
            Ff = [mc[i] * Sf[i] - np.cross(hc[i], Sn[i]) for i in range(nb)]
            bias = np.array([Sn[i] @ fcn[i] + Sf[i] @ fcf[i] for i in range(nb)])
            M = np.zeros((nb, nb))   # M[r, j] = F_r . S_j for the descendants r of j (and symmetric), 0 between different branches
            for j in range(nb):
                for r in desc[j]:
                    M[r, j] = M[j, r] = Fn[r] @ Sn[j] + Ff[r] @ Sf[j]
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
                # LDL^T in REVERSED body order (virtual index nb - 1 - body: leaves first), as the kernel runs it
                perm = np.arange(nb)[::-1]
                Hv = (M + np.diag(dimp))[np.ix_(perm, perm)]
                yv = (tau - bias)[perm]
                L, D = np.eye(nb), np.zeros(nb)
                Hw = Hv.copy()
                for kk in range(nb):
                    D[kk] = Hw[kk, kk]
                    L[kk + 1:, kk] = Hw[kk + 1:, kk] / D[kk]
                    Hw[kk + 1:, kk + 1:] -= np.outer(L[kk + 1:, kk], Hw[kk, kk + 1:])
                z = np.linalg.solve(L, yv) / D
                qdd = np.linalg.solve(L.T, z)[perm]           # back to body order (perm is its own inverse)
                pivots = D[perm]
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
    return (q, qd, pivots) if want_pivots else (q, qd)


def aba_diagonals(m, p, q, tgt_dimp):
    """D_j of the articulated-body recursion (world frame, spatial quantities about the world origin) for the configuration q:
    what the contact solve uses as joint compliance -- computed independently of the LDL^T above (6x6 articulated inertias)."""
    nb = m.nb
    parent = [m.parent[i] for i in range(nb)]
    bq, bp = np.array(m.base_quat[:], float), np.array(m.base_pos[:], float)
    Rw, ow = [None] * nb, [None] * nb
    S = [None] * nb
    IA = [None] * nb
    for i in range(nb):
        Rp, op = (q2R(bq), bp) if parent[i] < 0 else (Rw[parent[i]], ow[parent[i]])
        Rt = Rp @ np.array(m.tree_R[i][:], float).reshape(3, 3)
        oi = op + Rp @ np.array(m.tree_p[i][:], float)
        ax = Rt[:, 2]
        if m.jtype[i] == 0:
            c, s = np.cos(q[i]), np.sin(q[i])
            Rw[i] = Rt @ np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]]); ow[i] = oi
            S[i] = np.concatenate([ax, np.cross(oi, ax)])
        else:
            Rw[i] = Rt; ow[i] = oi + q[i] * ax
            S[i] = np.concatenate([np.zeros(3), ax])
        mass = m.mass[i]
        com = np.array(m.mcom[i][:], float) / mass if mass > 0 else np.zeros(3)
        Io = sym(np.array(m.inertia[i][:], float))
        Icm = Io - mass * (com @ com * np.eye(3) - np.outer(com, com))
        cw = ow[i] + Rw[i] @ com
        cx = np.array([[0, -cw[2], cw[1]], [cw[2], 0, -cw[0]], [-cw[1], cw[0], 0]])
        I6 = np.zeros((6, 6))     # spatial inertia about the world origin, [angular; linear] ordering
        I6[:3, :3] = Rw[i] @ Icm @ Rw[i].T - mass * cx @ cx
        I6[:3, 3:] = mass * cx
        I6[3:, :3] = -mass * cx
        I6[3:, 3:] = mass * np.eye(3)
        IA[i] = I6
    D = np.zeros(nb)
    for i in range(nb - 1, -1, -1):
        U = IA[i] @ S[i]
        D[i] = S[i] @ U + tgt_dimp[i]
        if parent[i] >= 0:
            IA[parent[i]] = IA[parent[i]] + IA[i] - np.outer(U, U) / D[i]
    return D


def check(names=("panda_gripper", "omnipanda", "albert", "jackal", "boxer"), K=3, T=6, verbose=True):
    from oracle import oracle as orc
    from scenes import robot_setup
    links = {"albert": "mmrobot_link7", "omnipanda": "panda_hand", "jackal": "ee_link", "boxer": "ee_link", "panda_gripper": "panda_hand"}
    worst_all = 0.0
    for name in names:
        sc, p, state0 = robot_setup(name, links[name], K=K, T=T, u_lim=0.4)
        rng = np.random.default_rng(4)
        actions = rng.uniform(-0.4, 0.4, (T, sc.nu, K)).astype(np.float32)
        actions[:, :, 0] *= 6.0                                  # one rollout far beyond the effort limits -> saturation re-solve
        st_ref, _ = orc.rollout(sc.model, p, state0, actions, use_double=True, root0=sc.root_state0.astype(np.float32))
        nb = sc.model.nb
        worst = 0.0
        for k in range(K):
            q, qd, piv = rollout(sc.model, p, state0, actions[:, :, k].astype(float), want_pivots=True)
            worst = max(worst, np.abs(q - st_ref[:nb, k]).max(), np.abs(qd - st_ref[nb:2 * nb, k]).max() * 1e-2)
        # the leaves-first pivots ARE the articulated-body diagonals D_j (joint compliance of the contact solve); unsaturated drive terms
        h = p.dt / p.substeps
        dimp = np.array([sc.model.armature[i] + h * (sc.model.kd[i] + sc.model.damping[i]) for i in range(nb)])
        q0 = np.array(state0[:nb], float)
        _, _, piv0 = rollout(sc.model, _one_substep(p), state0, np.zeros((1, sc.nu)), want_pivots=True)
        D = aba_diagonals(sc.model, p, q0, dimp)
        dpiv = np.abs(piv0 / D - 1.0).max()
        if verbose:
            print(f"{name:14s} nb = {nb:2d}: max |q - oracle| (and 1e-2 |qd - oracle|) over {K} rollouts, T = {T}: {worst:.3e};  max |pivot / D_j - 1| = {dpiv:.2e}")
        assert worst < 5e-5, (name, worst)
        assert dpiv < 1e-6, (name, dpiv)      # (the model block holds tree_R and tree_quat as separately rounded float32)
        worst_all = max(worst_all, worst)
    return worst_all


def _one_substep(p):
    import copy
    p1 = copy.copy(p)
    p1.T, p1.substeps, p1.dt = 1, 1, p.dt / p.substeps
    return p1




# ------------------------------------------------------------------------------------------------------------------------------
# The contact solve of rollout_team.cu in GENERALISED COORDINATES, for the part that is new against the oracle: a free body's three
# linear velocity components (world) and three angular velocity components IN BODY AXES, one scalar inverse inertia each, contact rows
# cached once per substep.  Case: a free box on the ground plane (up to 8 corner contacts), one model step of the heijn_push scene with
# the robot far away -- the oracle solves it with world-frame inverse inertia tensors and point velocities.
def free_box_step(m, p, box_state, f=0):
    """One model step of free body f alone on the ground (no other contact partner in reach); box_state = its 13 root-state numbers.
    Returns the 13 numbers after the step."""
    h = p.dt / p.substeps
    x, qt, v, w = (np.array(box_state[0:3], float), np.array(box_state[3:7], float), np.array(box_state[7:10], float),
                   np.array(box_state[10:13], float))
    shape = next(s for s in range(m.nshapes) if m.shape_owner_kind[s] == 2 and m.shape_owner[s] == f)
    half = np.array(m.free_half[f][:], float)
    mass = float(m.free_mass[f])
    Iinv = 1.0 / (mass / 3.0 * np.array([half[1] ** 2 + half[2] ** 2, half[0] ** 2 + half[2] ** 2, half[0] ** 2 + half[1] ** 2]))
    minv = np.concatenate([np.full(3, 1.0 / mass), Iinv])                      # one scalar inverse inertia per coordinate
    kp, kd = float(m.contact_kp), float(m.contact_kd)
    gamma, beta = 1.0 / (h * (h * kp + kd)), h * kp / (h * kp + kd)
    mu = 0.5 * (float(m.shape_friction[shape]) + float(m.ground_friction))
    g = np.array(m.gravity[:], float)
    for _ in range(p.substeps):
        R = q2R(qt)
        sq, sp = np.array(m.shape_quat[shape][:], float), np.array(m.shape_pos[shape][:], float)
        Rs, cs = R @ q2R(sq), x + R @ sp
        contacts = []
        for ix in (-1, 1):
            for iy in (-1, 1):
                for iz in (-1, 1):
                    pt = Rs @ (np.array([ix, iy, iz]) * half) + cs
                    if pt[2] < float(m.ground_margin):
                        contacts.append((pt, -pt[2]))
        if m.free_gravity[f]:
            v = v + h * g
        u = np.concatenate([v, R.T @ w])                                       # generalised velocities: world linear, BODY-axis angular
        n = np.array([0.0, 0.0, 1.0])
        e = np.array([1.0, 0, 0]) if abs(n[0]) < 0.9 else np.array([0, 1.0, 0])
        t1 = np.cross(n, e); t1 /= np.linalg.norm(t1); t2 = np.cross(n, t1)
        rows, consts, lam = [], [], []
        for pt, d in contacts:                                                 # once per contact: rows, inverse effective masses, bias
            r = pt - x
            J = [np.concatenate([dr, R.T @ np.cross(r, dr)]) for dr in (n, t1, t2)]
            k = [float((Jr * Jr) @ minv) for Jr in J]
            bias = min(beta * d / h, float(m.max_depen)) if d > 0 else d / h
            rows.append(J)
            consts.append((bias, 1.0 / (k[0] + gamma) if k[0] > 1e-9 else 0.0, 1.0 / k[1] if k[1] > 1e-9 else 0.0, 1.0 / k[2] if k[2] > 1e-9 else 0.0))
            lam.append(np.zeros(3))
        for _it in range(m.contact_iters):
            for c, J in enumerate(rows):
                bias, ikn, ikt1, ikt2 = consts[c]
                if not ikn > 0:
                    continue
                vn, v1, v2 = (float(Jr @ u) for Jr in J)
                ln, lt1, lt2 = lam[c]
                ln_new = max(0.0, ln + (-vn + bias - gamma * ln) * ikn)
                lim = mu * ln_new
                lt1_new = min(max(lt1 - v1 * ikt1, -lim), lim) if ikt1 > 0 else lt1
                lt2_new = min(max(lt2 - v2 * ikt2, -lim), lim) if ikt2 > 0 else lt2
                u = u + minv * (J[0] * (ln_new - ln) + J[1] * (lt1_new - lt1) + J[2] * (lt2_new - lt2))
                lam[c] = np.array([ln_new, lt1_new, lt2_new])
        v, w = u[:3], R @ u[3:]
        x = x + h * v
        wq = np.array([w[0], w[1], w[2], 0.0])
        qt = qt + 0.5 * h * qmul(wq, qt)
        qt = qt / np.linalg.norm(qt)
    return np.concatenate([x, qt, v, w])


def check_free_box(verbose=True):
    from oracle import oracle as orc
    from scenes import push_setup
    sc, p, s0 = push_setup(K=4, T=2, obstacles=False, block_pos=(3.0, -2.0, 0.1), robot_pos=(0.0, 1.5, 0.05))
    m = sc.model
    nd2 = 2 * sc.ndof
    NS = orc.lib().oracle_state_size(__import__("ctypes").byref(m))
    rng = np.random.default_rng(12)
    worst = 0.0
    for trial in range(6):
        ang = rng.uniform(-0.04, 0.04, 3)                                      # a slightly tilted box, some corners in the ground, some above it
        qt = np.array([np.sin(ang[0] / 2), 0, 0, np.cos(ang[0] / 2)])
        qt = qmul(np.array([0, np.sin(ang[1] / 2), 0, np.cos(ang[1] / 2)]), qt)
        qt = qmul(np.array([0, 0, np.sin(0.7 + ang[2]), np.cos(0.7 + ang[2])]), qt)
        half = np.array(m.free_half[0][:], float)
        box = np.concatenate([[3.0, -2.0, half[2] - 0.002 * trial], qt, rng.uniform(-0.5, 0.5, 3) * [1, 1, 0.2], rng.uniform(-1.0, 1.0, 3)])
        state = np.zeros((NS, p.K), np.float32)
        state[:nd2] = s0[:, None]
        state[nd2:nd2 + 13] = box.astype(np.float32)[:, None]
        box32 = state[nd2:nd2 + 13, 0].astype(float)                          # (the state rows are float32: start both from the same numbers)
        ref, _ = orc.rollout(m, p, None, np.zeros((p.T, sc.nu, p.K), np.float32), 0, 1, state=state.copy(), root0=sc.root_state0, want_obs=False, use_double=True)
        mine = free_box_step(m, p, box32)
        err = np.abs(mine - ref[nd2:nd2 + 13, 0]).max()
        moved = np.abs(ref[nd2 + 7:nd2 + 13, 0] - box32[7:13]).max()
        assert moved > 1e-2, "the contacts must have done something"
        worst = max(worst, err)
    if verbose:
        print(f"free box on the ground, generalised-coordinate Gauss-Seidel vs the oracle's world-frame solve, 6 poses: max |state difference| = {worst:.2e}")
    assert worst < 5e-6, worst
    return worst


if __name__ == "__main__":
    check()
    check_free_box()
