"""MuJoCo-shaped model -> `irlosc_model` (include/irlosc.h) for the fused state provider.

The reference asks MuJoCo for M, J, qfrc_bias and the EE poses every timestep
(robot.py:68-72, device.py:93-95,115-143, osc.py:191).  The fused step computes
them on the GPU from (q, dq), so the library needs the rigid-body description
MuJoCo holds: body frames, hinge axes, inertial parameters.  This module
reduces a model that exposes the `mujoco_py` arrays (`body_parentid`,
`body_pos`, `body_quat`, `body_jntadr/jntnum`, `jnt_axis`, `jnt_pos`,
inertials) to one entry per JOINT: bodies without joints are folded into the
nearest ancestor body that has one - frames composed, inertias lumped with the
parallel-axis theorem - and bodies welded to the world are dropped (they do not
move).  Pure host-side setup, run once per controller.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import _native


def _quat_to_mat(q) -> np.ndarray:
    w, x, y, z = [float(v) for v in q]
    s = 2.0 / (w * w + x * x + y * y + z * z)
    xs, ys, zs = x * s, y * s, z * s
    return np.array([[1.0 - (y * ys + z * zs), x * ys - w * zs, x * zs + w * ys],
                     [x * ys + w * zs, 1.0 - (x * xs + z * zs), y * zs - w * xs],
                     [x * zs - w * ys, y * zs + w * xs, 1.0 - (x * xs + y * ys)]])


def _mat_to_quat(R: np.ndarray) -> np.ndarray:
    K = np.array([[R[0, 0] - R[1, 1] - R[2, 2], 0, 0, 0],
                  [R[0, 1] + R[1, 0], R[1, 1] - R[0, 0] - R[2, 2], 0, 0],
                  [R[0, 2] + R[2, 0], R[1, 2] + R[2, 1], R[2, 2] - R[0, 0] - R[1, 1], 0],
                  [R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1], R[0, 0] + R[1, 1] + R[2, 2]]]) / 3.0
    vals, vecs = np.linalg.eigh(K)
    q = vecs[[3, 0, 1, 2], np.argmax(vals)]
    return -q if q[0] < 0 else q


class _Carrier:
    """Frame of a body relative to the joint-carrying body it is welded to."""

    def __init__(self, model):
        self.model = model
        nb = getattr(model, "n_robot_bodies", model.nbody)
        self.joint_of_body = {}
        for b in range(nb):
            if int(model.body_jntnum[b]) > 0:
                if int(model.body_jntnum[b]) != 1:
                    raise ValueError("body %d carries %d joints; the fused step supports one hinge per body"
                                     % (b, int(model.body_jntnum[b])))
                self.joint_of_body[b] = int(model.body_jntadr[b])

    def resolve(self, body: int):
        """(carrier body or 0 for the world, R, p): pose of `body` in its carrier's frame."""
        m = self.model
        R, p = np.eye(3), np.zeros(3)
        b = int(body)
        while b != 0 and b not in self.joint_of_body:
            Rb, pb = _quat_to_mat(m.body_quat[b]), np.asarray(m.body_pos[b], dtype=np.float64)
            R, p = Rb @ R, Rb @ p + pb
            b = int(m.body_parentid[b])
        return b, R, p


def reduce_model(model, joint_ids_all: Sequence[int], ee_bodies: Sequence[str],
                 ft_sites: Sequence[Optional[str]], gravity=(0.0, 0.0, -9.81)) -> "_native.Model":
    """Build the `irlosc_model` for robot joints `joint_ids_all` (robot-local order) and the target
    devices' EE bodies / F-T sites (target order, None = no sensor)."""
    car = _Carrier(model)
    local = {int(g): i for i, g in enumerate(joint_ids_all)}
    n = len(joint_ids_all)
    if n > _native.MAX_N:
        raise ValueError("at most %d joints" % _native.MAX_N)
    out = _native.Model()
    out.n_joints = n
    for i in range(3):
        out.gravity[i] = float(gravity[i])
    nb = getattr(model, "n_robot_bodies", model.nbody)
    # lumped inertial parameters per carrier body
    parts = {b: [] for b in car.joint_of_body}
    for b in range(1, nb):
        it = model.body_inertial[b]
        if it is None:
            continue
        c, R, p = car.resolve(b)
        if c == 0:
            continue                       # welded to the world
        ipos, iquat, mass, diag = it
        Ri = R @ _quat_to_mat(iquat)
        parts[c].append((float(mass), p + R @ np.asarray(ipos, dtype=np.float64),
                         Ri @ np.diag(np.asarray(diag, dtype=np.float64)) @ Ri.T))
    for g, i in local.items():
        b = int(model.jnt_bodyid[g])
        if np.any(np.asarray(model.jnt_pos[g]) != 0.0):
            raise ValueError("joint %d is not anchored at its body origin" % g)
        jm = out.joint[i]
        # frame relative to the parent JOINT body: own (pos, quat) composed with welded ancestors
        pb = int(model.body_parentid[b])
        c, R, p = car.resolve(pb)
        Rb, posb = _quat_to_mat(model.body_quat[b]), np.asarray(model.body_pos[b], dtype=np.float64)
        Rt, pt = R @ Rb, R @ posb + p
        jm.parent = local.get(car.joint_of_body.get(c, -1), -1) if c != 0 else -1
        q = _mat_to_quat(Rt)
        for k in range(3):
            jm.pos[k] = float(pt[k])
            jm.axis[k] = float(model.jnt_axis[g][k])
        for k in range(4):
            jm.quat[k] = float(q[k])
        ps = parts[b]
        mass = sum(m for m, _, _ in ps)
        com = sum(m * c_ for m, c_, _ in ps) / mass if mass > 0 else np.zeros(3)
        I = np.zeros((3, 3))
        for m, c_, Ic in ps:
            d = c_ - com
            I += Ic + m * (d @ d * np.eye(3) - np.outer(d, d))
        jm.mass = float(mass)
        for k in range(3):
            jm.com[k] = float(com[k])
        for k, (r_, c_) in enumerate([(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]):
            jm.inertia[k] = float(I[r_, c_])
    for d in range(_native.MAX_DEVICES):
        out.ee[d].joint = -1
        out.ft[d].joint = -1
        out.ee[d].quat[0] = out.ft[d].quat[0] = 1.0
    for d, name in enumerate(ee_bodies):
        c, R, p = car.resolve(model.body_name2id(name))
        fr = out.ee[d]
        fr.joint = local[car.joint_of_body[c]]
        q = _mat_to_quat(R)
        for k in range(3):
            fr.pos[k] = float(p[k])
        for k in range(4):
            fr.quat[k] = float(q[k])
    for d, name in enumerate(ft_sites):
        if name is None:
            continue
        s = model.site_name2id(name)
        c, R, p = car.resolve(int(model.site_bodyid[s]))
        Rs = R @ _quat_to_mat(model.site_quat[s])
        ps = R @ np.asarray(model.site_pos[s], dtype=np.float64) + p
        fr = out.ft[d]
        fr.joint = local[car.joint_of_body[c]]
        q = _mat_to_quat(Rs)
        for k in range(3):
            fr.pos[k] = float(ps[k])
        for k in range(4):
            fr.quat[k] = float(q[k])
    return out


def model_for_layout(model, robot_joint_ids_all: Sequence[int], layout) -> "_native.Model":
    """`reduce_model` with the EE bodies / F-T sites of `layout`'s target devices."""
    sites = set(getattr(model, "site_names", []))
    ee = [d.ee_body for d in layout.devices]
    if any(not e for e in ee):
        raise ValueError("layout lacks EE body names (compile_layout fills them from Device.EE)")
    ft = [("ft_frame_" + d.name) if ("ft_frame_" + d.name) in sites else None for d in layout.devices]
    return reduce_model(model, robot_joint_ids_all, ee, ft)
