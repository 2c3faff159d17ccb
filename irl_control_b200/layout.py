"""Flattening of `Device` / `Robot` / `OSC` configuration into index tables.

Everything the reference resolves in its constructors - joint chains and
actuator maps (`device.py:41-74`), the robot joint set (`robot.py:26-32`), the
gain vectors (`osc.py:35-39`) - plus the target order chosen by the caller of
`generate` (`osc.py:136-138,156`) is reduced to one small, immutable
description.  It is handed to the CUDA library as `irlosc_params`
(include/irlosc.h) and, as a plain dict, to the test oracle.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native


@dataclass(frozen=True)
class DeviceLayout:
    name: str
    ctrlr_dof: tuple            # 6 bools, xyz then abg (device.py:36)
    joint_ids_all: tuple        # robot-local (robot.py:64 uses the global ids as local ones)
    actuator_trnids: tuple      # robot-local joints returned by generate (osc.py:207)
    ctrl_idxs: tuple            # sim.data.ctrl slots those forces go to (osc.py:208)
    dx_idx: tuple               # Robot J_idxs[name] in SUB-DEVICE order (robot.py:52-55)
    has_max_vel: bool
    max_vel: tuple
    kp: float
    kv: float
    ko: float
    k: tuple
    d: tuple
    gain_vectors: Optional[tuple] = None   # (task_space_gains[6], lamb[6]) as OSC.__init__ stored them (osc.py:35-39);
                                           # None: derived from kp / kv / ko
    ee_joint: int = -1          # robot-local id of the last joint of Device.joint_ids (device.py:62-64)
    ee_body: str = ""           # Device.EE, the body whose pose / Jacobian the device reads (device.py:93-95,125)

    @property
    def n_rows(self) -> int:
        return int(sum(bool(x) for x in self.ctrlr_dof))


@dataclass(frozen=True)
class OscLayout:
    n: int
    devices: tuple              # DeviceLayout, TARGET order
    use_g: bool = True
    admittance: bool = False
    nullspace_kv: Optional[float] = None
    joint_parent: Optional[tuple] = None    # robot-local parent joint per joint (-1 root); None = unknown
    check_topology: bool = False

    @property
    def D(self) -> int:
        return len(self.devices)

    @property
    def k(self) -> int:
        return sum(d.n_rows for d in self.devices)

    @property
    def n_ctrl(self) -> int:
        return sum(len(d.actuator_trnids) for d in self.devices)

    @property
    def ctrl_slices(self) -> List[slice]:
        out, at = [], 0
        for d in self.devices:
            out.append(slice(at, at + len(d.actuator_trnids)))
            at += len(d.actuator_trnids)
        return out

    @property
    def tile_entries(self) -> int:
        """Doubles per instance in the batch-interleaved tile layout (csrc/osc_lane.cuh build_tile_spec;
        `irlosc_tile_entries` is the authority, the GPU tests compare the two); 0 when the controller has none."""
        arms = [d for d in self.devices if d.name != "base" and d.ee_joint in (6, 18)]
        base = [d for d in self.devices if d.ee_joint == 0]
        if self.joint_parent is None or len(arms) != 2 or len(arms) + len(base) != self.D or len(base) > 1:
            return 0
        kd = arms[0].n_rows
        if arms[1].n_rows != kd or kd not in (3, 6) or (base and base[0].n_rows != 1):
            return 0
        dev_n = 31 if self.admittance else 16          # poses 14 + max_vel 2 [+ F/T frame 9 + raw wrench 6]
        return 2 + (dev_n if base else 0) + 2 * (35 + 2 * 31 + 7 * kd + 6 + dev_n)

    def as_dict(self) -> Dict:
        return {
            "n": self.n, "use_g": self.use_g, "admittance": self.admittance,
            "nullspace_kv": self.nullspace_kv,
            "joint_parent": None if self.joint_parent is None else list(self.joint_parent),
            "devices": [{
                "name": d.name, "ctrlr_dof": list(d.ctrlr_dof), "joint_ids_all": list(d.joint_ids_all),
                "actuator_trnids": list(d.actuator_trnids), "ctrl_idxs": list(d.ctrl_idxs),
                "dx_idx": list(d.dx_idx), "has_max_vel": d.has_max_vel, "max_vel": list(d.max_vel),
                "kp": d.kp, "kv": d.kv, "ko": d.ko, "k": list(d.k), "d": list(d.d),
                "gain_vectors": None if d.gain_vectors is None else [list(d.gain_vectors[0]), list(d.gain_vectors[1])],
                "ee_joint": d.ee_joint,
            } for d in self.devices],
        }

    def to_c_params(self) -> "_native.Params":
        if self.D > _native.MAX_DEVICES:
            raise ValueError("at most %d target devices are supported" % _native.MAX_DEVICES)
        p = _native.Params()
        p.abi_version = _native.ABI_VERSION
        p.n = self.n
        p.n_devices = self.D
        p.use_g = int(self.use_g)
        p.admittance = int(self.admittance)
        p.has_nullspace = int(self.nullspace_kv is not None)
        p.nullspace_kv = float(self.nullspace_kv or 0.0)
        p.has_topology = int(self.joint_parent is not None)
        p.check_topology = int(self.check_topology)
        for j in range(_native.MAX_N):
            p.joint_parent[j] = -1
        if self.joint_parent is not None:
            for j, v in enumerate(self.joint_parent):
                p.joint_parent[j] = int(v)
        for i, d in enumerate(self.devices):
            c = p.dev[i]
            for j in range(6):
                c.ctrlr_dof[j] = int(bool(d.ctrlr_dof[j]))
            c.n_joints_all = len(d.joint_ids_all)
            for j, v in enumerate(d.joint_ids_all):
                c.joint_ids_all[j] = int(v)
            c.n_ctrl = len(d.actuator_trnids)
            for j, v in enumerate(d.actuator_trnids):
                c.actuator_trnids[j] = int(v)
            for j, v in enumerate(d.dx_idx):
                c.dx_idx[j] = int(v)
            c.has_max_vel = int(d.has_max_vel)
            c.max_vel[0], c.max_vel[1] = (float(d.max_vel[0]), float(d.max_vel[1])) if d.has_max_vel else (0.0, 0.0)
            c.kp, c.kv, c.ko = float(d.kp), float(d.kv), float(d.ko)
            for j in range(3):
                c.k[j] = float(d.k[j])
                c.d[j] = float(d.d[j])
            c.has_gain_vectors = int(d.gain_vectors is not None)
            if d.gain_vectors is not None:
                for j in range(6):
                    c.task_space_gains[j] = float(d.gain_vectors[0][j])
                    c.lamb[j] = float(d.gain_vectors[1][j])
            c.ee_joint = int(d.ee_joint)
        return p


def joint_parents(model, joint_ids_all) -> Optional[tuple]:
    """Robot-local parent joint of every robot joint (-1 for roots), from the same body tree that
    `Device.__init__` walks (device.py:41-64).  None when the model lacks `jnt_bodyid`."""
    if not hasattr(model, "jnt_bodyid"):
        return None
    local = {int(g): i for i, g in enumerate(joint_ids_all)}
    out = []
    for g in joint_ids_all:
        g = int(g)
        body = int(model.jnt_bodyid[g])
        parent = -1
        if g > int(model.body_jntadr[body]):           # several joints on one body: chained
            parent = g - 1
        else:
            b = int(model.body_parentid[body])
            while b != 0 and int(model.body_jntnum[b]) == 0:
                b = int(model.body_parentid[b])
            if b != 0:
                parent = int(model.body_jntadr[b]) + int(model.body_jntnum[b]) - 1
        out.append(local.get(parent, -1))
    return tuple(out)


def compile_layout(robot, device_configs: Dict[str, Dict], target_names: Sequence[str],
                   nullspace_config: Optional[Dict], use_g: bool, admittance: bool,
                   check_topology: bool = False) -> OscLayout:
    """Build the layout for one target order.

    robot          : a `Robot` (ours or anything with sub_devices_dict / joint_ids_all)
    device_configs : device name -> controller config dict with kp, kv, ko, k, d
    """
    local = {int(g): i for i, g in enumerate(robot.joint_ids_all)}
    # J_idxs as Robot.jacobians numbers them: every sub-device, construction order
    j_idxs, row = {}, 0
    for name, dev in robot.sub_devices_dict.items():
        rows = int(np.sum(dev.ctrlr_dof))
        j_idxs[name] = tuple(range(row, row + rows))
        row += rows
    devs = []
    for name in target_names:
        dev = robot.sub_devices_dict[name]          # KeyError for unknown names, as the reference
        cfg = device_configs[name]
        mv = dev.max_vel
        devs.append(DeviceLayout(
            name=name,
            ctrlr_dof=tuple(bool(x) for x in dev.ctrlr_dof),
            joint_ids_all=tuple(local[int(g)] for g in dev.joint_ids_all),
            actuator_trnids=tuple(local[int(g)] for g in dev.actuator_trnids),
            ctrl_idxs=tuple(int(x) for x in dev.ctrl_idxs),
            dx_idx=j_idxs[name],
            has_max_vel=mv is not None,
            max_vel=(float(mv[0]), float(mv[1])) if mv is not None else (0.0, 0.0),
            kp=float(cfg['kp']), kv=float(cfg['kv']), ko=float(cfg['ko']),
            k=tuple(float(x) for x in cfg['k']), d=tuple(float(x) for x in cfg['d']),
            gain_vectors=((tuple(float(x) for x in cfg['task_space_gains']), tuple(float(x) for x in cfg['lamb']))
                          if ('task_space_gains' in cfg and 'lamb' in cfg) else None),
            ee_joint=local.get(int(dev.joint_ids[-1]), -1) if len(dev.joint_ids) else -1,
            ee_body=str(getattr(dev, 'EE', '')),
        ))
    return OscLayout(
        n=int(robot.num_joints_total), devices=tuple(devs), use_g=bool(use_g),
        admittance=bool(admittance),
        nullspace_kv=None if nullspace_config is None else float(nullspace_config['kv']),
        joint_parent=joint_parents(robot.sim.model, robot.joint_ids_all),
        check_topology=bool(check_topology))


# ---------------------------------------------------------------- MuJoCo's sparse inertia (IRLOSC_M_QM)
def qm_index(joint_parent: Sequence[int]) -> Tuple[np.ndarray, np.ndarray]:
    """(rows, cols) of the entries of `mjData.qM` in storage order for a dof tree given by `dof_parentid`:
    dof i owns M[i, i], M[i, parent(i)], M[i, parent(parent(i))], ... down to its root, dofs in order
    (the walk of `mj_fullM`, which robot.py:69 calls).  `M[:, rows, cols]` is the robot's part of qM."""
    rows, cols = [], []
    for i in range(len(joint_parent)):
        j = i
        while j >= 0:
            if joint_parent[j] >= j:
                raise ValueError("joint_parent[%d]=%d does not precede it" % (j, joint_parent[j]))
            rows.append(i)
            cols.append(j)
            j = joint_parent[j]
    return np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64)


def qm_size(joint_parent: Sequence[int]) -> int:
    """nM of the tree: 155 for the DualUR5's 25 dofs."""
    return int(len(qm_index(joint_parent)[0]))
