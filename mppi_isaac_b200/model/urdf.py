"""URDF -> rollout-model compiler (host side, numpy only).

Replaces the IsaacGym asset importer used at
``mppiisaac/utils/isaacgym_utils.py:14-28`` (``gym.load_asset`` with
``fix_base_link`` / ``disable_gravity``) on the rollout path: it reads a URDF,
derives missing inertials from the collision geometry at the importer's default
density (all panda URDFs carry zero ``<inertial>`` tags),
merges fixed-joint links into their moving ancestor and emits the flat constant
block (``MppibModel`` in ``include/mppib.h``) that both the CUDA rollout kernel
and the CPU oracle consume.

Frame conventions
-----------------
* ``R`` of a transform has the child axes as columns (maps child coords to
  parent coords); ``p`` is the child origin in parent coords.
* URDF ``rpy`` is fixed-axis roll/pitch/yaw: ``R = Rz(yaw) Ry(pitch) Rx(roll)``.
* Every moving body gets a *body frame* equal to its URDF link frame rotated by
  a constant ``A`` such that the joint axis becomes ``+z``; kernels therefore
  only know revolute-z and prismatic-z joints.
"""
from __future__ import annotations

import json
import math
import os
import struct
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

DEFAULT_DENSITY = 1000.0  # kg/m^3, importer default (reference never sets it: isaacgym_utils.py:15-22)


# ----------------------------------------------------------------------------------------
# small math helpers (float64 on the host; the blob is float32)
# ----------------------------------------------------------------------------------------
def rpy_to_R(rpy) -> np.ndarray:
    r, p, y = (float(v) for v in rpy)
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def quat_xyzw_to_R(q) -> np.ndarray:
    x, y, z, w = (float(v) for v in q)
    n = math.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / n, y / n, z / n, w / n
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
        ]
    )


def R_to_quat_xyzw(R: np.ndarray) -> np.ndarray:
    """Branch-on-largest-diagonal conversion; the kernels use the same branches."""
    m00, m11, m22 = R[0, 0], R[1, 1], R[2, 2]
    tr = m00 + m11 + m22
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        w = 0.25 * s
        x = (R[2, 1] - R[1, 2]) / s
        y = (R[0, 2] - R[2, 0]) / s
        z = (R[1, 0] - R[0, 1]) / s
    elif m00 > m11 and m00 > m22:
        s = math.sqrt(1.0 + m00 - m11 - m22) * 2
        w = (R[2, 1] - R[1, 2]) / s
        x = 0.25 * s
        y = (R[0, 1] + R[1, 0]) / s
        z = (R[0, 2] + R[2, 0]) / s
    elif m11 > m22:
        s = math.sqrt(1.0 + m11 - m00 - m22) * 2
        w = (R[0, 2] - R[2, 0]) / s
        x = (R[0, 1] + R[1, 0]) / s
        y = 0.25 * s
        z = (R[1, 2] + R[2, 1]) / s
    else:
        s = math.sqrt(1.0 + m22 - m00 - m11) * 2
        w = (R[1, 0] - R[0, 1]) / s
        x = (R[0, 2] + R[2, 0]) / s
        y = (R[1, 2] + R[2, 1]) / s
        z = 0.25 * s
    return np.array([x, y, z, w])


def skew(v) -> np.ndarray:
    x, y, z = v
    return np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]], dtype=float)


def axis_to_z_rotation(axis: np.ndarray) -> np.ndarray:
    """Rotation A with A @ [0,0,1] = axis (minimal rotation)."""
    a = np.asarray(axis, dtype=float)
    a = a / np.linalg.norm(a)
    z = np.array([0.0, 0.0, 1.0])
    c = float(a @ z)
    if c > 1 - 1e-12:
        return np.eye(3)
    if c < -1 + 1e-12:
        return np.diag([1.0, -1.0, -1.0])  # pi about x
    v = np.cross(z, a)
    s = np.linalg.norm(v)
    vx = skew(v)
    return np.eye(3) + vx + vx @ vx * ((1 - c) / (s * s))


def _snap(M: np.ndarray, eps: float = 1e-9) -> np.ndarray:
    """Snap 1.57079632679-style rounding residue to exact 0 / +-1."""
    M = np.array(M, dtype=float)
    M[np.abs(M) < eps] = 0.0
    M[np.abs(M - 1) < eps] = 1.0
    M[np.abs(M + 1) < eps] = -1.0
    return M


# ----------------------------------------------------------------------------------------
# inertia of primitives / meshes in the geometry's own frame: (mass, com, I_com)
# ----------------------------------------------------------------------------------------
def box_inertia(size, density=DEFAULT_DENSITY):
    x, y, z = (float(v) for v in size)
    m = density * x * y * z
    I = np.diag([m * (y * y + z * z) / 12, m * (x * x + z * z) / 12, m * (x * x + y * y) / 12])
    return m, np.zeros(3), I


def sphere_inertia(radius, density=DEFAULT_DENSITY):
    r = float(radius)
    m = density * 4.0 / 3.0 * math.pi * r**3
    return m, np.zeros(3), np.eye(3) * (0.4 * m * r * r)


def cylinder_inertia(radius, length, density=DEFAULT_DENSITY):
    r, l = float(radius), float(length)
    m = density * math.pi * r * r * l
    Ixx = m * (3 * r * r + l * l) / 12
    return m, np.zeros(3), np.diag([Ixx, Ixx, 0.5 * m * r * r])


def load_mesh(path: str) -> Tuple[np.ndarray, np.ndarray]:
    """Minimal OBJ / binary-or-ascii STL reader -> (vertices (N,3), triangles (M,3))."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".obj":
        verts, tris = [], []
        with open(path, "r") as f:
            for line in f:
                if line.startswith("v "):
                    verts.append([float(t) for t in line.split()[1:4]])
                elif line.startswith("f "):
                    idx = [int(tok.split("/")[0]) for tok in line.split()[1:]]
                    idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
                    for k in range(1, len(idx) - 1):
                        tris.append([idx[0], idx[k], idx[k + 1]])
        return np.asarray(verts, dtype=float), np.asarray(tris, dtype=np.int64)
    if ext == ".stl":
        with open(path, "rb") as f:
            data = f.read()
        ntri = struct.unpack_from("<I", data, 80)[0] if len(data) >= 84 else 0
        if 84 + 50 * ntri == len(data):
            arr = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=ntri, offset=84)
            verts = arr["v"].reshape(-1, 3).astype(float)
            return verts, np.arange(3 * ntri).reshape(-1, 3)
        verts = []
        for line in data.decode("ascii", "ignore").splitlines():
            t = line.split()
            if len(t) == 4 and t[0] == "vertex":
                verts.append([float(v) for v in t[1:]])
        verts = np.asarray(verts, dtype=float)
        return verts, np.arange(len(verts)).reshape(-1, 3)
    raise NotImplementedError(f"mesh format {ext} not supported: {path}")


def mesh_inertia(verts: np.ndarray, tris: np.ndarray, density=DEFAULT_DENSITY, scale=(1, 1, 1)):
    """Exact mass properties of a closed triangle mesh by signed tetrahedra (origin apex)."""
    v = verts * np.asarray(scale, dtype=float)
    a, b, c = v[tris[:, 0]], v[tris[:, 1]], v[tris[:, 2]]
    det = np.einsum("ij,ij->i", a, np.cross(b, c))  # 6 * signed tetra volume
    vol = det.sum() / 6.0
    if vol < 0:  # inward facing winding
        det, vol = -det, -vol
    com = (det[:, None] * (a + b + c) / 4.0).sum(0) / 6.0 / vol
    # second moments  int x_i x_j dV over each tetra (origin,a,b,c):  det/120 * (sum_pairs)
    S = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            t = (
                2 * (a[:, i] * a[:, j] + b[:, i] * b[:, j] + c[:, i] * c[:, j])
                + a[:, i] * b[:, j] + a[:, j] * b[:, i]
                + a[:, i] * c[:, j] + a[:, j] * c[:, i]
                + b[:, i] * c[:, j] + b[:, j] * c[:, i]
            )
            S[i, j] = (det * t).sum() / 120.0
    S *= density
    m = density * vol
    I_o = np.eye(3) * np.trace(S) - S  # about the mesh origin
    I_c = I_o - m * (np.eye(3) * (com @ com) - np.outer(com, com))
    return m, com, I_c


# ----------------------------------------------------------------------------------------
# URDF parsing
# ----------------------------------------------------------------------------------------
@dataclass
class UGeom:
    kind: str  # box | sphere | cylinder | mesh
    size: List[float]
    R: np.ndarray
    p: np.ndarray
    mesh: Optional[str] = None
    scale: Tuple[float, float, float] = (1.0, 1.0, 1.0)


@dataclass
class ULink:
    name: str
    inertial: Optional[Tuple[float, np.ndarray, np.ndarray]]  # (m, com, I_com) in the link frame
    collisions: List[UGeom] = field(default_factory=list)


@dataclass
class UJoint:
    name: str
    jtype: str
    parent: str
    child: str
    R: np.ndarray
    p: np.ndarray
    axis: np.ndarray
    lower: float = -1e30
    upper: float = 1e30
    effort: float = 1e30
    velocity: float = 1e30
    damping: float = 0.0


def _floats(s, n=None, default=None):
    if s is None:
        return list(default)
    v = [float(t) for t in s.replace(",", " ").split()]
    return v


def _origin(elem):
    o = elem.find("origin") if elem is not None else None
    if o is None:
        return np.eye(3), np.zeros(3)
    return rpy_to_R(_floats(o.get("rpy"), default=(0, 0, 0))), np.asarray(_floats(o.get("xyz"), default=(0, 0, 0)), float)


def resolve_mesh(filename: str, urdf_path: str) -> str:
    """``package://pkg/rest`` -> <assets/urdf>/pkg/rest ; relative paths are URDF relative."""
    urdf_dir = os.path.dirname(os.path.abspath(urdf_path))
    if filename.startswith("package://"):
        rest = filename[len("package://"):]
        d = urdf_dir
        for _ in range(6):
            cand = os.path.join(d, rest)
            if os.path.exists(cand):
                return cand
            d = os.path.dirname(d)
        return os.path.join(urdf_dir, rest)
    if os.path.isabs(filename):
        return filename
    return os.path.join(urdf_dir, filename)


def parse_urdf(path: str) -> Tuple[Dict[str, ULink], List[UJoint]]:
    root = ET.parse(path).getroot()
    links: Dict[str, ULink] = {}
    for le in root.findall("link"):
        inertial = None
        ie = le.find("inertial")
        if ie is not None and ie.find("mass") is not None:
            m = float(ie.find("mass").get("value"))
            R, p = _origin(ie)
            ine = ie.find("inertia")
            g = lambda k: float(ine.get(k, 0.0)) if ine is not None else 0.0
            I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
            inertial = (m, p, R @ I @ R.T)
        cols = []
        for ce in le.findall("collision"):
            R, p = _origin(ce)
            ge = ce.find("geometry")
            if ge is None or len(ge) == 0:
                continue
            g0 = ge[0]
            if g0.tag == "box":
                cols.append(UGeom("box", _floats(g0.get("size")), R, p))
            elif g0.tag == "sphere":
                cols.append(UGeom("sphere", [float(g0.get("radius"))], R, p))
            elif g0.tag == "cylinder":
                cols.append(UGeom("cylinder", [float(g0.get("radius")), float(g0.get("length"))], R, p))
            elif g0.tag == "mesh":
                sc = tuple(_floats(g0.get("scale"), default=(1, 1, 1)))
                cols.append(UGeom("mesh", [], R, p, mesh=g0.get("filename"), scale=sc))
        links[le.get("name")] = ULink(le.get("name"), inertial, cols)
    joints: List[UJoint] = []
    for je in root.findall("joint"):
        if je.find("parent") is None or je.find("child") is None:
            continue
        R, p = _origin(je)
        ax = je.find("axis")
        axis = np.asarray(_floats(ax.get("xyz")) if ax is not None else [1, 0, 0], float)
        j = UJoint(je.get("name"), je.get("type"), je.find("parent").get("link"), je.find("child").get("link"), R, p, axis)
        lim = je.find("limit")
        if lim is not None:
            j.lower = float(lim.get("lower", -1e30))
            j.upper = float(lim.get("upper", 1e30))
            j.effort = float(lim.get("effort", 1e30))
            j.velocity = float(lim.get("velocity", 1e30))
        if j.jtype == "continuous":
            j.lower, j.upper = -1e30, 1e30
        dyn = je.find("dynamics")
        if dyn is not None:
            j.damping = float(dyn.get("damping", 0.0))
        joints.append(j)
    return links, joints


def link_mass_properties(link: ULink, urdf_path: str, density=DEFAULT_DENSITY):
    """(m, com, I_com) in the link frame; collision-derived when no <inertial> is given."""
    if link.inertial is not None:
        return link.inertial
    m_tot, first, parts = 0.0, np.zeros(3), []
    for g in link.collisions:
        if g.kind == "box":
            m, c, I = box_inertia(g.size, density)
        elif g.kind == "sphere":
            m, c, I = sphere_inertia(g.size[0], density)
        elif g.kind == "cylinder":
            m, c, I = cylinder_inertia(g.size[0], g.size[1], density)
        else:
            v, t = load_mesh(resolve_mesh(g.mesh, urdf_path))
            m, c, I = mesh_inertia(v, t, density, g.scale)
        c_l = g.R @ c + g.p
        I_l = g.R @ I @ g.R.T
        parts.append((m, c_l, I_l))
        m_tot += m
        first += m * c_l
    if m_tot <= 0:
        return 0.0, np.zeros(3), np.zeros((3, 3))
    com = first / m_tot
    I = np.zeros((3, 3))
    for m, c, Ic in parts:
        d = c - com
        I += Ic + m * (np.eye(3) * (d @ d) - np.outer(d, d))
    return m_tot, com, I


# ----------------------------------------------------------------------------------------
# compiled robot
# ----------------------------------------------------------------------------------------
@dataclass
class RobotModel:
    """Flat articulated model: one moving body per 1-DoF joint, fixed links merged."""
    name: str
    fixed_base: bool
    dof_names: List[str]
    link_names: List[str]
    parent: List[int]
    jtype: List[int]
    tree_R: np.ndarray      # (nb,3,3)
    tree_p: np.ndarray      # (nb,3)
    mass: np.ndarray        # (nb,)
    mcom: np.ndarray        # (nb,3)
    inertia_o: np.ndarray   # (nb,3,3) about body origin
    q_lo: np.ndarray
    q_hi: np.ndarray
    qd_max: np.ndarray
    effort: np.ndarray
    damping: np.ndarray
    link_body: List[int]
    link_R: np.ndarray      # (nl,3,3) link frame in owning-body frame
    link_p: np.ndarray      # (nl,3)
    link_collisions: List[List[dict]] = field(default_factory=list)  # per link, primitives in link frame
    planar_base: bool = False   # bodies 0,1,2 are virtual joints (world x, world y, yaw) of a floating base reduced to the plane

    @property
    def nb(self) -> int:
        return len(self.parent)

    @property
    def nlinks(self) -> int:
        return len(self.link_names)

    def to_json(self) -> dict:
        d = {}
        for k, v in self.__dict__.items():
            d[k] = v.tolist() if isinstance(v, np.ndarray) else v
        return d

    @staticmethod
    def from_json(d: dict) -> "RobotModel":
        arr = ("tree_R", "tree_p", "mass", "mcom", "inertia_o", "q_lo", "q_hi", "qd_max", "effort", "damping", "link_R", "link_p")
        kw = {k: (np.asarray(v, dtype=float) if k in arr else v) for k, v in d.items()}
        return RobotModel(**kw)


_JT = {"revolute": 0, "continuous": 0, "prismatic": 1}


def compile_urdf(urdf_path: str, fixed_base: bool = True, density: float = DEFAULT_DENSITY,
                 root_mass_override: Optional[float] = None) -> RobotModel:
    """Compile a URDF into a RobotModel (fixed-base articulations; see module docstring)."""
    links, joints = parse_urdf(urdf_path)
    children: Dict[str, List[UJoint]] = {n: [] for n in links}
    is_child = set()
    for j in joints:
        if j.parent in links and j.child in links:
            children[j.parent].append(j)
            is_child.add(j.child)
    roots = [n for n in links if n not in is_child]

    def subtree(n):
        return 1 + sum(subtree(j.child) for j in children[n])

    # an orphan second root (franka_panda_gripper.urdf:168-192 "panda_link8") is ignored
    root = max(roots, key=subtree)

    link_names: List[str] = []
    link_body: List[int] = []
    link_T: List[Tuple[np.ndarray, np.ndarray]] = []
    dof_names: List[str] = []
    body = dict(parent=[], jtype=[], tree_R=[], tree_p=[], q_lo=[], q_hi=[], qd_max=[], effort=[], damping=[])
    # accumulated mass properties per body, expressed in body frame about the body origin
    acc_m: List[float] = []
    acc_h: List[np.ndarray] = []
    acc_I: List[np.ndarray] = []
    link_cols: List[List[dict]] = []

    def add_link(name: str, owner: int, R_bl: np.ndarray, p_bl: np.ndarray):
        """Register link `name` whose frame sits at (R_bl, p_bl) in body `owner`'s frame."""
        link_names.append(name)
        link_body.append(owner)
        link_T.append((R_bl, p_bl))
        lk = links[name]
        cols = []
        for g in lk.collisions:
            d = dict(kind=g.kind, size=list(g.size), R=g.R.tolist(), p=g.p.tolist(), mesh=g.mesh)
            if g.kind == "mesh":      # axis-aligned bounding box in the geometry frame (collision proxy of the rollout path)
                v, _ = load_mesh(resolve_mesh(g.mesh, urdf_path))
                v = v * np.asarray(g.scale, float)
                d["aabb_center"] = (0.5 * (v.min(0) + v.max(0))).tolist()
                d["aabb_half"] = (0.5 * (v.max(0) - v.min(0))).tolist()
            cols.append(d)
        link_cols.append(cols)
        m, c, Ic = link_mass_properties(lk, urdf_path, density)
        if name == root and root_mass_override is not None and m > 0:
            # isaacgym_wrapper.py:450-456 overwrites body-0 mass (inertia left untouched)
            m = float(root_mass_override)
        if owner >= 0 and m > 0:
            cb = R_bl @ c + p_bl
            Ib = R_bl @ Ic @ R_bl.T + m * (np.eye(3) * (cb @ cb) - np.outer(cb, cb))
            acc_m[owner] += m
            acc_h[owner] += m * cb
            acc_I[owner] += Ib
        for j in children[name]:
            if j.jtype == "fixed":
                add_link(j.child, owner, R_bl @ j.R, p_bl + R_bl @ j.p)
            elif j.jtype in _JT:
                A = axis_to_z_rotation(j.axis)
                R_t = _snap(R_bl @ j.R @ A)
                p_t = p_bl + R_bl @ j.p
                idx = len(body["parent"])
                body["parent"].append(owner)
                body["jtype"].append(_JT[j.jtype])
                body["tree_R"].append(R_t)
                body["tree_p"].append(p_t)
                body["q_lo"].append(j.lower)
                body["q_hi"].append(j.upper)
                body["qd_max"].append(j.velocity)
                body["effort"].append(j.effort)
                body["damping"].append(j.damping)
                dof_names.append(j.name)
                acc_m.append(0.0)
                acc_h.append(np.zeros(3))
                acc_I.append(np.zeros((3, 3)))
                add_link(j.child, idx, _snap(A.T), np.zeros(3))
            else:
                raise NotImplementedError(f"joint type {j.jtype} ({j.name}) not supported")

    if fixed_base:
        add_link(root, -1, np.eye(3), np.zeros(3))
    else:
        # floating base reduced to the plane: three virtual joints (world x, world y, yaw) carry the root link
        Ax, Ay = axis_to_z_rotation(np.array([1.0, 0, 0])), axis_to_z_rotation(np.array([0, 1.0, 0]))
        for name, jt, Rt, par in (("__base_x", 1, Ax, -1), ("__base_y", 1, Ax.T @ Ay, 0), ("__base_yaw", 0, Ay.T, 1)):
            body["parent"].append(par); body["jtype"].append(jt); body["tree_R"].append(_snap(Rt)); body["tree_p"].append(np.zeros(3))
            body["q_lo"].append(-1e30); body["q_hi"].append(1e30); body["qd_max"].append(1e30); body["effort"].append(1e30); body["damping"].append(0.0)
            dof_names.append(name)
            acc_m.append(0.0); acc_h.append(np.zeros(3)); acc_I.append(np.zeros((3, 3)))
        add_link(root, 2, np.eye(3), np.zeros(3))

    nb = len(body["parent"])
    return RobotModel(
        name=os.path.splitext(os.path.basename(urdf_path))[0],
        fixed_base=fixed_base,
        dof_names=dof_names,
        link_names=link_names,
        parent=body["parent"],
        jtype=body["jtype"],
        tree_R=np.asarray(body["tree_R"], float).reshape(nb, 3, 3),
        tree_p=np.asarray(body["tree_p"], float).reshape(nb, 3),
        mass=np.asarray(acc_m, float),
        mcom=np.asarray(acc_h, float).reshape(nb, 3),
        inertia_o=np.asarray(acc_I, float).reshape(nb, 3, 3),
        q_lo=np.asarray(body["q_lo"], float),
        q_hi=np.asarray(body["q_hi"], float),
        qd_max=np.asarray(body["qd_max"], float),
        effort=np.asarray(body["effort"], float),
        damping=np.asarray(body["damping"], float),
        link_body=link_body,
        link_R=np.asarray([t[0] for t in link_T], float).reshape(-1, 3, 3),
        link_p=np.asarray([t[1] for t in link_T], float).reshape(-1, 3),
        link_collisions=link_cols,
        planar_base=not fixed_base,
    )


# ----------------------------------------------------------------------------------------
# host-side forward kinematics (float64) -- used by tests and by the facade for K=1 queries
# ----------------------------------------------------------------------------------------
def forward_kinematics(model: RobotModel, q, base_pos=(0, 0, 0), base_quat=(0, 0, 0, 1)):
    """World poses of every link: (positions (nl,3), quats xyzw (nl,4))."""
    q = np.asarray(q, float)
    Rb, pb = quat_xyzw_to_R(base_quat), np.asarray(base_pos, float)
    Rw, pw = [], []
    for i in range(model.nb):
        Rp, pp = (Rb, pb) if model.parent[i] < 0 else (Rw[model.parent[i]], pw[model.parent[i]])
        R, p = model.tree_R[i], model.tree_p[i]
        if model.jtype[i] == 0:
            c, s = math.cos(q[i]), math.sin(q[i])
            Rj, pj = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]]), np.zeros(3)
        else:
            Rj, pj = np.eye(3), np.array([0, 0, q[i]])
        Rw.append(Rp @ R @ Rj)
        pw.append(pp + Rp @ (p + R @ pj))
    pos, quat = [], []
    for l in range(model.nlinks):
        b = model.link_body[l]
        Rp, pp = (Rb, pb) if b < 0 else (Rw[b], pw[b])
        pos.append(pp + Rp @ model.link_p[l])
        quat.append(R_to_quat_xyzw(Rp @ model.link_R[l]))
    return np.asarray(pos), np.asarray(quat)


def save_compiled(model: RobotModel, path: str):
    with open(path, "w") as f:
        json.dump(model.to_json(), f, indent=1)


def load_compiled(path: str) -> RobotModel:
    with open(path) as f:
        return RobotModel.from_json(json.load(f))
