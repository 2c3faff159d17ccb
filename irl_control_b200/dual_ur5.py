"""DualUR5 rigid-body model: topology tables + batched kinematics/dynamics.

What the reference obtains from MuJoCo for the OSC path, provided without
MuJoCo (absent from this image) for (a) constructing `Device`/`Robot` index
maps exactly like `device.py:41-74` does against a `mujoco_py` model, and
(b) producing physically consistent synthetic states for parity tests and
the benchmark:

    mj_fullM        (robot.py:69)        -> `Dynamics.M`       CRBA-equivalent
    mj_jacBody      (device.py:125-128)  -> `Dynamics.jacp/jacr`
    qfrc_bias       (osc.py:191)         -> `Dynamics.bias`    RNEA(q, dq, 0)
    xpos / xquat    (device.py:93-95)    -> `Dynamics.xpos/xquat`
    site_xmat       (device.py:140)      -> `Dynamics.site_xmat`

This is input synthesis, not the control law: nothing here is on the timed
path.  Tensors are torch float64 so the same code runs on the host (tests)
and on the GPU (benchmark input generation).  Contacts, equality
constraints, armature and damping are not modelled.  Bodies without an <inertial> element
take mass / centre of mass / inertia from their mesh geom at MuJoCo's default density, as MuJoCo's compiler does
(`base_link_ur5right / _ur5left`, welded to the stand: tools/extract_dual_ur5.py, restated from MuJoCo 2.0 / 2.1's
mesh algorithm - MuJoCo itself is absent, so this is unpinned); bodies with neither are massless.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import dual_ur5_model as _tables

GRAVITY = (0.0, 0.0, -9.81)


class DualUR5Model:
    """`mujoco_py`-model-shaped view of the DualUR5 tree.

    Only the attributes `Device.__init__` reads (device.py:41-74) plus what
    `MujocoApp.set_free_joint_qpos` needs.  `n_free_objects` appends free
    bodies after the robot, as the admit_test / insertion scenes do
    (`admit_test_scene.xml:8-15`), so that `nv` > 25 while robot DoF ids
    stay 0..24.
    """

    def __init__(self, n_free_objects: int = 0):
        names = ["world"]
        parent = [0]
        pos = [(0.0, 0.0, 0.0)]
        quat = [(1.0, 0.0, 0.0, 0.0)]
        inertial = [None]
        jntadr, jntnum = [-1], [0]
        self.joint_names: List[str] = []
        jnt_body, jnt_axis, jnt_pos, jnt_range = [], [], [], []
        self.site_names: List[str] = []
        site_body, site_pos, site_quat = [], [], []
        index = {"world": 0}
        for (bname, bparent, bpos, bquat, binert, bjoints, bsites) in _tables.BODIES:
            bid = len(names)
            index[bname] = bid
            names.append(bname)
            parent.append(index[bparent])
            pos.append(bpos)
            quat.append(bquat)
            inertial.append(binert)
            jntadr.append(len(self.joint_names) if bjoints else -1)
            jntnum.append(len(bjoints))
            for (jname, jaxis, jpos, jrange) in bjoints:
                self.joint_names.append(jname)
                jnt_body.append(bid)
                jnt_axis.append(jaxis)
                jnt_pos.append(jpos)
                jnt_range.append(jrange if jrange else (-np.inf, np.inf))
            for (sname, spos, squat) in bsites:
                self.site_names.append(sname)
                site_body.append(bid)
                site_pos.append(spos)
                site_quat.append(squat)
        self.n_robot_bodies = len(names)
        self.nv_robot = len(self.joint_names)
        # free objects appended after the robot: 6 dof / 7 qpos each
        self.n_free_objects = int(n_free_objects)
        self.free_joint_names = ["free_joint_%d" % i for i in range(self.n_free_objects)]
        for i in range(self.n_free_objects):
            names.append("free_object_%d" % i)
            index[names[-1]] = len(names) - 1
            parent.append(0)
            pos.append((0.5 + 0.1 * i, 0.5, 0.0))
            quat.append((1.0, 0.0, 0.0, 0.0))
            inertial.append(((0.0, 0.0, 0.0), (1.0, 0.0, 0.0, 0.0), 0.1, (1e-4, 1e-4, 1e-4)))
            jntadr.append(self.nv_robot + i)
            jntnum.append(1)

        self.body_names = names
        self._body_index = index
        self.body_parentid = np.asarray(parent, dtype=np.int32)
        self.body_pos = np.asarray(pos, dtype=np.float64)
        self.body_quat = np.asarray(quat, dtype=np.float64)
        self.body_inertial = inertial
        self.body_jntadr = np.asarray(jntadr, dtype=np.int32)
        self.body_jntnum = np.asarray(jntnum, dtype=np.int32)
        self.jnt_bodyid = np.asarray(jnt_body, dtype=np.int32)
        self.jnt_axis = np.asarray(jnt_axis, dtype=np.float64)
        self.jnt_pos = np.asarray(jnt_pos, dtype=np.float64)
        self.jnt_range = np.asarray(jnt_range, dtype=np.float64)
        self.site_bodyid = np.asarray(site_body, dtype=np.int32)
        self.site_pos = np.asarray(site_pos, dtype=np.float64)
        self.site_quat = np.asarray(site_quat, dtype=np.float64)
        self.nbody = len(names)
        self.njnt = self.nv_robot + self.n_free_objects
        self.nv = self.nv_robot + 6 * self.n_free_objects
        self.nq = self.nv_robot + 7 * self.n_free_objects
        # qpos address of every joint (hinges first, then 7 per free joint)
        self.jnt_qposadr = np.asarray(
            list(range(self.nv_robot))
            + [self.nv_robot + 7 * i for i in range(self.n_free_objects)], dtype=np.int32)
        jid = {n: i for i, n in enumerate(self.joint_names)}
        self.actuator_names = [a[1] for a in _tables.ACTUATORS]
        self.actuator_kind = [a[0] for a in _tables.ACTUATORS]
        self.actuator_trnid = np.asarray(
            [[jid[a[2]], -1] for a in _tables.ACTUATORS], dtype=np.int32)
        self.nu = len(self.actuator_names)
        self.sensor_names = [s[1] for s in _tables.SENSORS]
        self.sensor_site = [self.site_names.index(s[2]) for s in _tables.SENSORS]
        self.nsensordata = 3 * len(self.sensor_names)

    # --- the mujoco_py accessors Device.__init__ uses -------------------
    def body_name2id(self, name: str) -> int:
        if name not in self._body_index:
            raise ValueError('No "body" with name %s exists.' % name)
        return self._body_index[name]

    def joint_id2name(self, jid: int) -> str:
        jid = int(jid)
        if jid < self.nv_robot:
            return self.joint_names[jid]
        return self.free_joint_names[jid - self.nv_robot]

    def joint_name2id(self, name: str) -> int:
        if name in self.joint_names:
            return self.joint_names.index(name)
        if name in self.free_joint_names:
            return self.nv_robot + self.free_joint_names.index(name)
        raise ValueError('No "joint" with name %s exists.' % name)

    def site_name2id(self, name: str) -> int:
        if name not in self.site_names:
            raise ValueError('No "site" with name %s exists.' % name)
        return self.site_names.index(name)


# ----------------------------------------------------------------------
# batched rigid-body math (torch, float64)
# ----------------------------------------------------------------------

def _quat_to_mat(q: torch.Tensor) -> torch.Tensor:
    w, x, y, z = q.unbind(-1)
    s = 2.0 / (w * w + x * x + y * y + z * z)
    xs, ys, zs = x * s, y * s, z * s
    r = torch.stack([
        1.0 - (y * ys + z * zs), x * ys - w * zs, x * zs + w * ys,
        x * ys + w * zs, 1.0 - (x * xs + z * zs), y * zs - w * xs,
        x * zs - w * ys, y * zs + w * xs, 1.0 - (x * xs + y * ys)], dim=-1)
    return r.reshape(q.shape[:-1] + (3, 3))


def _mat_to_quat(r: torch.Tensor) -> torch.Tensor:
    """Rotation matrix -> unit quaternion (w >= 0 branch-free via 4 candidates)."""
    m00, m01, m02 = r[..., 0, 0], r[..., 0, 1], r[..., 0, 2]
    m10, m11, m12 = r[..., 1, 0], r[..., 1, 1], r[..., 1, 2]
    m20, m21, m22 = r[..., 2, 0], r[..., 2, 1], r[..., 2, 2]
    cand = torch.stack([
        torch.stack([1 + m00 + m11 + m22, m21 - m12, m02 - m20, m10 - m01], -1),
        torch.stack([m21 - m12, 1 + m00 - m11 - m22, m01 + m10, m02 + m20], -1),
        torch.stack([m02 - m20, m01 + m10, 1 - m00 + m11 - m22, m12 + m21], -1),
        torch.stack([m10 - m01, m02 + m20, m12 + m21, 1 - m00 - m11 + m22], -1),
    ], dim=-2)
    diag = torch.stack([cand[..., i, i] for i in range(4)], dim=-1)
    best = diag.argmax(dim=-1)
    q = torch.gather(cand, -2, best[..., None, None].expand(best.shape + (1, 4))).squeeze(-2)
    q = q / q.norm(dim=-1, keepdim=True)
    return torch.where(q[..., :1] < 0, -q, q)


def _axis_angle_mat(axis: torch.Tensor, angle: torch.Tensor) -> torch.Tensor:
    """Rodrigues; axis (3,) unit, angle (B,) -> (B,3,3)."""
    x, y, z = axis.tolist()
    c, s = torch.cos(angle), torch.sin(angle)
    t = 1.0 - c
    r = torch.stack([
        t * x * x + c, t * x * y - s * z, t * x * z + s * y,
        t * x * y + s * z, t * y * y + c, t * y * z - s * x,
        t * x * z - s * y, t * y * z + s * x, t * z * z + c], dim=-1)
    return r.reshape(angle.shape + (3, 3))


@dataclass
class Dynamics:
    """Per-instance quantities for a batch of (q, dq); robot DoF only (n = 25)."""
    xpos: torch.Tensor        # (B, nbody, 3)   body frame origins
    xmat: torch.Tensor        # (B, nbody, 3, 3)
    xquat: torch.Tensor       # (B, nbody, 4)
    axis_w: torch.Tensor      # (B, n, 3)       joint axes, world
    anchor_w: torch.Tensor    # (B, n, 3)       joint anchors, world
    M: torch.Tensor           # (B, n, n)
    bias: torch.Tensor        # (B, n)          RNEA(q, dq, 0) incl. gravity
    site_xmat: torch.Tensor   # (B, nsite, 3, 3)
    model: "DualUR5Model"

    def jac_body(self, body: int):
        """(jacp, jacr), each (B, 3, n): Jacobian of the body frame origin."""
        return _body_jacobian(self.model, self, body, self.xpos[:, body])


def _ancestor_mask(model: DualUR5Model) -> np.ndarray:
    """mask[b, j] = 1 if joint j moves body b."""
    n = model.nv_robot
    mask = np.zeros((model.n_robot_bodies, n), dtype=np.float64)
    for b in range(1, model.n_robot_bodies):
        mask[b] = mask[model.body_parentid[b]]
        if model.body_jntnum[b] > 0:
            for jj in range(model.body_jntnum[b]):
                mask[b, model.body_jntadr[b] + jj] = 1.0
    return mask


def _body_jacobian(model, dyn, body, point):
    mask = torch.as_tensor(_ancestor_mask(model)[body], dtype=point.dtype, device=point.device)
    jacr = dyn.axis_w * mask[None, :, None]                      # (B, n, 3)
    jacp = torch.cross(jacr, point[:, None, :] - dyn.anchor_w, dim=-1)
    return jacp.transpose(1, 2).contiguous(), jacr.transpose(1, 2).contiguous()


def dynamics(model: DualUR5Model, q: torch.Tensor, dq: torch.Tensor,
             ddq: Optional[torch.Tensor] = None, need_M: bool = True,
             gravity: Sequence[float] = GRAVITY) -> Dynamics:
    """Forward kinematics, joint-space inertia and bias forces for a batch.

    q, dq: (B, 25) float64.  `ddq` (optional) makes `bias` the full inverse
    dynamics RNEA(q, dq, ddq); tests use it to cross-check M column by column.
    """
    assert q.dtype == torch.float64 and q.shape[-1] == model.nv_robot
    B, n = q.shape
    dev = q.device
    f64 = dict(dtype=torch.float64, device=dev)
    nb = model.n_robot_bodies
    T = lambda a: torch.as_tensor(np.asarray(a), **f64)

    xpos: List[torch.Tensor] = [torch.zeros(B, 3, **f64)]
    xmat: List[torch.Tensor] = [torch.eye(3, **f64).expand(B, 3, 3)]
    axis_w = [None] * n
    anchor_w = [None] * n
    for b in range(1, nb):
        p = model.body_parentid[b]
        R = xmat[p] @ _quat_to_mat(T(model.body_quat[b]))
        o = xpos[p] + (xmat[p] @ T(model.body_pos[b]))
        if model.body_jntnum[b] > 0:
            assert model.body_jntnum[b] == 1
            j = int(model.body_jntadr[b])
            ax = T(model.jnt_axis[j])
            ax = ax / ax.norm()
            anchor = o + R @ T(model.jnt_pos[j])
            R = R @ _axis_angle_mat(ax, q[:, j])
            o = anchor - R @ T(model.jnt_pos[j])
            axis_w[j] = R @ ax
            anchor_w[j] = anchor
        xpos.append(o)
        xmat.append(R)
    xpos_t = torch.stack(xpos, dim=1)
    xmat_t = torch.stack(xmat, dim=1)
    axis_t = torch.stack(axis_w, dim=1)
    anchor_t = torch.stack(anchor_w, dim=1)
    site_xmat = torch.stack([
        xmat_t[:, model.site_bodyid[s]] @ _quat_to_mat(T(model.site_quat[s]))
        for s in range(len(model.site_names))], dim=1)

    dyn = Dynamics(xpos=xpos_t, xmat=xmat_t, xquat=_mat_to_quat(xmat_t), axis_w=axis_t,
                   anchor_w=anchor_t, M=None, bias=None, site_xmat=site_xmat, model=model)

    # ---- inertial frames ------------------------------------------------
    com, Iw, mass = {}, {}, {}
    for b in range(1, nb):
        it = model.body_inertial[b]
        if it is None:
            continue
        ipos, iquat, m, diag = it
        Ri = xmat_t[:, b] @ _quat_to_mat(T(iquat))
        com[b] = xpos_t[:, b] + xmat_t[:, b] @ T(ipos)
        Iw[b] = (Ri * T(diag)[None, None, :]) @ Ri.transpose(1, 2)
        mass[b] = float(m)

    # ---- joint-space inertia: sum_b m Jc^T Jc + Jw^T I Jw ---------------
    if need_M:
        M = torch.zeros(B, n, n, **f64)
        for b in com:
            jp, jr = _body_jacobian(model, dyn, b, com[b])
            M += mass[b] * (jp.transpose(1, 2) @ jp)
            M += jr.transpose(1, 2) @ (Iw[b] @ jr)
        dyn.M = 0.5 * (M + M.transpose(1, 2))

    # ---- recursive Newton-Euler, world frame, moments about the origin --
    g = T(gravity)
    zero3 = torch.zeros(B, 3, **f64)
    w = [zero3] * nb        # angular velocity
    dw = [zero3] * nb       # angular acceleration
    acc_o = [(-g).expand(B, 3)] + [None] * (nb - 1)   # linear acc. of body origin
    cross = lambda a, b_: torch.cross(a, b_, dim=-1)
    for b in range(1, nb):
        p = model.body_parentid[b]
        if model.body_jntnum[b] > 0:
            j = int(model.body_jntadr[b])
            a = axis_t[:, j]
            c = anchor_t[:, j]
            # anchor as a point fixed in the parent
            rc = c - xpos_t[:, p]
            acc_c = acc_o[p] + cross(dw[p], rc) + cross(w[p], cross(w[p], rc))
            w[b] = w[p] + a * dq[:, j:j + 1]
            dw[b] = dw[p] + cross(w[p], a) * dq[:, j:j + 1]
            if ddq is not None:
                dw[b] = dw[b] + a * ddq[:, j:j + 1]
            ro = xpos_t[:, b] - c
            acc_o[b] = acc_c + cross(dw[b], ro) + cross(w[b], cross(w[b], ro))
        else:
            w[b], dw[b] = w[p], dw[p]
            ro = xpos_t[:, b] - xpos_t[:, p]
            acc_o[b] = acc_o[p] + cross(dw[p], ro) + cross(w[p], cross(w[p], ro))
    F = [zero3.clone() for _ in range(nb)]
    N0 = [zero3.clone() for _ in range(nb)]
    for b in com:
        rc = com[b] - xpos_t[:, b]
        a_c = acc_o[b] + cross(dw[b], rc) + cross(w[b], cross(w[b], rc))
        f = mass[b] * a_c
        Iwb = (Iw[b] @ w[b][..., None])[..., 0]
        nloc = (Iw[b] @ dw[b][..., None])[..., 0] + cross(w[b], Iwb)
        F[b] = F[b] + f
        N0[b] = N0[b] + nloc + cross(com[b], f)
    tau = torch.zeros(B, n, **f64)
    for b in range(nb - 1, 0, -1):
        if model.body_jntnum[b] > 0:
            j = int(model.body_jntadr[b])
            tau[:, j] = (axis_t[:, j] * (N0[b] - cross(anchor_t[:, j], F[b]))).sum(-1)
        p = model.body_parentid[b]
        F[p] = F[p] + F[b]
        N0[p] = N0[p] + N0[b]
    dyn.bias = tau
    return dyn


# ----------------------------------------------------------------------
# seeded synthetic joint states (SURVEY.md section 8d "synthetic inputs")
# ----------------------------------------------------------------------

ARM_JOINTS = {"ur5right": list(range(1, 7)), "ur5left": list(range(13, 19))}
GRIPPER_JOINTS = {"ur5right": list(range(7, 13)), "ur5left": list(range(19, 25))}


def sample_joint_states(B: int, seed: int = 0):
    """q ~ U(-pi, pi) on stand + arm joints, U(0, 0.8) on gripper joints, dq ~ N(0, 0.3^2)."""
    rng = np.random.default_rng(seed)
    q = rng.uniform(-np.pi, np.pi, size=(B, 25))
    for name in GRIPPER_JOINTS:
        idx = GRIPPER_JOINTS[name]
        q[:, idx] = rng.uniform(0.0, 0.8, size=(B, len(idx)))
    dq = rng.normal(0.0, 0.3, size=(B, 25))
    return q, dq
