"""Independent float64 forward dynamics from the Lagrangian (torch autograd) -- pins the oracle's ABA.

Nothing here uses spatial (6-D) algebra: the kinetic energy is T = 1/2 sum_i (m |v_com|^2 + w^T R I_c R^T w)
with v_com / w from the geometric Jacobians of the chain, M(q) = d2T/dqd2, and the bias term from the
Euler-Lagrange equations  M qdd + (dM/dt) qd - dT/dq = tau + gravity.
"""
import math

import torch

from mppi_isaac_b200.model.urdf import RobotModel


def _rz(q):
    c, s = torch.cos(q), torch.sin(q)
    z, o = torch.zeros_like(q), torch.ones_like(q)
    return torch.stack([torch.stack([c, -s, z]), torch.stack([s, c, z]), torch.stack([z, z, o])])


def mass_matrix_and_potential(model: RobotModel, q, gravity, base_R=None, base_p=None):
    dt = torch.float64
    nb = model.nb
    base_R = torch.eye(3, dtype=dt) if base_R is None else base_R
    base_p = torch.zeros(3, dtype=dt) if base_p is None else base_p
    Rw, pw, axes = [], [], []
    for i in range(nb):
        Rp, pp = (base_R, base_p) if model.parent[i] < 0 else (Rw[model.parent[i]], pw[model.parent[i]])
        Rt = torch.tensor(model.tree_R[i], dtype=dt)
        pt = torch.tensor(model.tree_p[i], dtype=dt)
        if model.jtype[i] == 0:
            R = Rp @ Rt @ _rz(q[i])
            p = pp + Rp @ pt
        else:
            R = Rp @ Rt
            p = pp + Rp @ (pt + Rt[:, 2] * q[i])
        Rw.append(R); pw.append(p); axes.append(R[:, 2])
    M = torch.zeros((nb, nb), dtype=dt)
    V = torch.zeros((), dtype=dt)
    g = torch.tensor(gravity, dtype=dt)
    for i in range(nb):
        m = float(model.mass[i])
        if m <= 0:
            continue
        c_b = torch.tensor(model.mcom[i] / m, dtype=dt)
        Io = torch.tensor(model.inertia_o[i], dtype=dt)
        Ic = Io - m * ((c_b @ c_b) * torch.eye(3, dtype=dt) - torch.outer(c_b, c_b))
        com = pw[i] + Rw[i] @ c_b
        Jv = torch.zeros((3, nb), dtype=dt)
        Jw = torch.zeros((3, nb), dtype=dt)
        j = i
        while j >= 0:
            if model.jtype[j] == 0:
                Jw[:, j] = axes[j]
                Jv[:, j] = torch.linalg.cross(axes[j], com - pw[j])
            else:
                Jv[:, j] = axes[j]
            j = model.parent[j]
        Iw = Rw[i] @ Ic @ Rw[i].T
        M = M + m * Jv.T @ Jv + Jw.T @ Iw @ Jw
        V = V - m * (g @ com)
    return M, V


def forward_dynamics(model: RobotModel, q, qd, tau, dimp, gravity=(0.0, 0.0, 0.0)):
    """qdd solving (M + diag(dimp)) qdd = tau - bias(q, qd) - dV/dq  (float64 lists/arrays in, tensor out)."""
    q = torch.tensor(q, dtype=torch.float64, requires_grad=True)
    qd = torch.tensor(qd, dtype=torch.float64)
    tau = torch.tensor(tau, dtype=torch.float64)
    dimp = torch.tensor(dimp, dtype=torch.float64)

    def Mfun(qq):
        return mass_matrix_and_potential(model, qq, gravity)[0]

    def Vfun(qq):
        return mass_matrix_and_potential(model, qq, gravity)[1]

    M = Mfun(q)
    dM = torch.autograd.functional.jacobian(Mfun, q)           # (nb, nb, nb): dM[i,j]/dq_k
    Mdot = torch.einsum("ijk,k->ij", dM, qd)
    dT_dq = 0.5 * torch.einsum("ijk,i,j->k", dM, qd, qd)
    dV = torch.autograd.functional.jacobian(Vfun, q)
    rhs = tau - Mdot @ qd + dT_dq - dV
    return torch.linalg.solve(M.detach() + torch.diag(dimp), rhs), M.detach()
