"""Adapter for the official `mujoco` Python bindings (the reference's README points to
its successor repo that uses them).  The package itself never imports `mujoco` - it is absent from
this image, so what is here is exercised with a stand-in object that carries the same arrays
(tests/test_fused_host.py) and is **untested against a real MuJoCo build**.

Three pieces a caller needs to drive the fused step from a live `mujoco.MjModel / MjData`:

    view  = MjModelView(mj_model, name2id=...)        # the mujoco_py-shaped arrays `reduce_model` reads
    model = reduce_model(view, robot_joint_ids, ee_bodies, ft_sites, gravity=mj_model.opt.gravity)
    q, dq = joint_state(mj_data, view, robot_joint_ids)            # what robot.py:60-65 pulls
    scatter_ctrl(mj_data.ctrl, layout, ctrl_row)                   # gain_test.py:146-147

Hinge joints only (every robot joint of scenes/dual_ur5.xml is a hinge); free joints of scene objects
come after the robot's joints in qpos / qvel and are ignored.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np


class MjModelView:
    """`mujoco.MjModel` seen through the attribute names `rigid_model.reduce_model` uses
    (those of `mujoco_py`, which the reference is written against: device.py:41-74)."""

    def __init__(self, m, name2id: Optional[Callable[[str, str], int]] = None):
        self._m = m
        self.nbody = int(m.nbody)
        self.n_robot_bodies = int(m.nbody)
        self.body_parentid = np.asarray(m.body_parentid)
        self.body_pos = np.asarray(m.body_pos, dtype=np.float64)
        self.body_quat = np.asarray(m.body_quat, dtype=np.float64)
        self.body_jntadr = np.asarray(m.body_jntadr)
        self.body_jntnum = np.asarray(m.body_jntnum)
        self.jnt_bodyid = np.asarray(m.jnt_bodyid)
        self.jnt_axis = np.asarray(m.jnt_axis, dtype=np.float64)
        self.jnt_pos = np.asarray(m.jnt_pos, dtype=np.float64)
        self.jnt_qposadr = np.asarray(m.jnt_qposadr)
        self.jnt_dofadr = np.asarray(m.jnt_dofadr)
        self.site_bodyid = np.asarray(m.site_bodyid)
        self.site_pos = np.asarray(m.site_pos, dtype=np.float64)
        self.site_quat = np.asarray(m.site_quat, dtype=np.float64)
        mass = np.asarray(m.body_mass, dtype=np.float64)
        ipos = np.asarray(m.body_ipos, dtype=np.float64)
        iquat = np.asarray(m.body_iquat, dtype=np.float64)
        inertia = np.asarray(m.body_inertia, dtype=np.float64)
        # (pos, quat, mass, diaginertia) per body, None for massless bodies - the tuple layout of dual_ur5_model.py
        self.body_inertial = [None if mass[b] == 0.0 else (tuple(ipos[b]), tuple(iquat[b]), float(mass[b]), tuple(inertia[b]))
                              for b in range(self.nbody)]
        if name2id is None:
            def name2id(kind, name):                       # official bindings: m.body(name).id / m.site(name).id
                return int(getattr(m, kind)(name).id)
        self._name2id = name2id
        self.site_names = _NameSet(self, "site")

    def body_name2id(self, name: str) -> int:
        return self._name2id("body", name)

    def site_name2id(self, name: str) -> int:
        return self._name2id("site", name)


class _NameSet:
    """`name in view.site_names` without enumerating the model's names."""

    def __init__(self, view, kind):
        self._view, self._kind = view, kind

    def __contains__(self, name):
        try:
            self._view._name2id(self._kind, name)
            return True
        except Exception:
            return False

    def __iter__(self):
        return iter(())


def joint_state(data, view: MjModelView, joint_ids: Sequence[int]):
    """(q, dq) of the robot joints from `MjData.qpos / qvel` (robot.py:60-65 reads qvel by joint id; the
    addresses coincide for the robot because its hinges come first)."""
    qadr = view.jnt_qposadr[np.asarray(joint_ids)]
    vadr = view.jnt_dofadr[np.asarray(joint_ids)]
    return np.asarray(data.qpos, dtype=np.float64)[qadr].copy(), np.asarray(data.qvel, dtype=np.float64)[vadr].copy()


def scatter_ctrl(ctrl, layout, ctrl_row) -> None:
    """`sim.data.ctrl[force_idx] = force` for every target device (gain_test.py:146-147) from one packed row."""
    for sl, dl in zip(layout.ctrl_slices, layout.devices):
        ctrl[list(dl.ctrl_idxs)] = np.asarray(ctrl_row)[sl]


def sparse_inertia(m, data, layout) -> np.ndarray:
    """`MjData.qM` as `state["qM"]` (IRLOSC_M_QM) - the array robot.py:69 expands with `mj_fullM` - after checking
    that the addressing the kernel derives from `layout.joint_parent` is the model's own: the robot's n dofs are
    the scene's first n, `dof_parentid` agrees, `dof_Madr` is the running sum of the dofs' depths.  Returns the
    scene's whole qM (length nM); the kernel reads the leading robot part, `m_stride = nM`."""
    from .layout import qm_index
    n = layout.n
    parent = np.asarray(m.dof_parentid)[:n]
    if layout.joint_parent is None or list(parent) != list(layout.joint_parent):
        raise ValueError("the model's dof_parentid[:%d] is not the layout's kinematic tree" % n)
    rows, _ = qm_index(layout.joint_parent)
    madr = np.searchsorted(rows, np.arange(n))              # first entry of every dof
    if list(np.asarray(m.dof_Madr)[:n]) != list(madr):
        raise ValueError("the model's dof_Madr[:%d] is not the running sum of the tree's depths" % n)
    qM = np.asarray(data.qM, dtype=np.float64)
    if qM.shape[0] < len(rows):
        raise ValueError("qM has %d entries, the robot alone needs %d" % (qM.shape[0], len(rows)))
    return qM
