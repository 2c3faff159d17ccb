"""`Device` - one controllable kinematic chain of the robot (host side).

Mirrors the public surface of the reference's `irl_control/device.py:21-213`:
constructor `(device_yml, model, sim, use_sim)`, the index attributes callers
and `OSC` read (`joint_ids`, `gripper_ids`, `joint_ids_all`, `ctrl_idxs`,
`actuator_trnids`, `ctrlr_dof*`, `max_vel`, `EE`, `name`) and the state
getters.  `sim` is any object with the mujoco_py `sim.model` / `sim.data`
accessors listed in SURVEY.md 8c (a real `MjSim`, or `SyntheticSim`).

In the batched B200 path these per-step getters are NOT called: the index
maps computed here are flattened once into the C-ABI parameter block
(`layout.py`) and state arrives as device-resident arrays.  The getters exist
so that single-robot callers of the reference API (`mujoco_app.py`-style
loops) keep working, one `generate()` at a time.
"""
import copy
from enum import Enum
from threading import Lock
from typing import Any, Callable, Dict

import numpy as np


class DeviceState(Enum):
    Q = 'Q'
    Q_ACTUATED = 'Q_ACTUATED'
    DQ = 'DQ'
    DQ_ACTUATED = 'DQ_ACTUATED'
    DDQ = 'DDQ'
    EE_XYZ = 'EE_XYZ'
    EE_XYZ_VEL = 'EE_XYZ_VEL'
    EE_QUAT = 'EE_QUAT'
    FORCE = 'FORCE'
    TORQUE = 'TORQUE'
    J = 'JACOBIAN'


# sensordata slices of the wrist F/T sensors (dual_ur5.xml:289-297 via device.py:150-167)
_FT_SLICES = {
    "ur5right": (slice(0, 3), slice(3, 6)),
    "ur5left": (slice(6, 9), slice(9, 12)),
}


def chain_joint_ids(model, ee_body: str, start_body_name=None):
    """Joint ids from the chain root to `ee_body` (device.py:41-64).

    Walks parents from the end-effector body and collects each visited body's
    joints; the walk ends at the first body whose parent is the world (id 0) or
    the optional `start_body`.  An unknown / absent start body means 0, as in the
    reference (bare `except`, device.py:41-44).
    """
    try:
        stop_at = model.body_name2id(start_body_name)
    except Exception:
        stop_at = 0
    ids, names = [], []
    body = model.body_name2id(ee_body)
    while True:
        up = model.body_parentid[body]
        if up == 0 or up == stop_at:
            break
        first = model.body_jntadr[body]
        here = [first + i for i in range(model.body_jntnum[body])]
        # prepend so the list ends up ordered root -> tip, joints of one body in model order
        ids = here + ids
        names = [model.joint_id2name(j) for j in here] + names
        body = up
    return np.array(ids, dtype=np.int64), names


class Device:
    def __init__(self, device_yml: Dict, model, sim, use_sim: bool):
        self.sim = sim
        self._use_sim = use_sim
        self.name = device_yml['name']
        self.max_vel = device_yml.get('max_vel')
        self.EE = device_yml['EE']
        self.ctrlr_dof_xyz = device_yml['ctrlr_dof_xyz']
        self.ctrlr_dof_abg = device_yml['ctrlr_dof_abg']
        # computed once: later edits of ctrlr_dof_abg do not change it
        self.ctrlr_dof = np.hstack([self.ctrlr_dof_xyz, self.ctrlr_dof_abg])
        self.start_angles = np.array(device_yml['start_angles'])
        self.num_gripper_joints = device_yml['num_gripper_joints']

        self.joint_ids, self.joint_names = chain_joint_ids(model, self.EE, device_yml.get('start_body'))
        first_gripper = self.joint_ids[-1] + 1
        self.gripper_ids = np.arange(first_gripper, first_gripper + self.num_gripper_joints)
        self.joint_ids_all = np.hstack([self.joint_ids, self.gripper_ids])

        # actuators whose transmission joint belongs to this device (device.py:72-74)
        trn = model.actuator_trnid[:, 0]
        self.ctrl_idxs = np.intersect1d(trn, self.joint_ids_all, return_indices=True)[1]
        self.actuator_trnids = trn[self.ctrl_idxs]

        # device.py:76-80 - same shape error as the reference when the chain length
        # does not match start_angles
        if self.name in ("ur5right", "ur5left", "base"):
            self.sim.data.qpos[self.joint_ids] = np.copy(self.start_angles)
        self.sim.forward()

        if np.sum(self.ctrlr_dof) > len(self.joint_ids):
            print("Fewer DOF than specified")

        d = self.sim.data
        self._getters: Dict[DeviceState, Callable[[], Any]] = {
            DeviceState.Q: lambda: d.qpos[self.joint_ids_all],
            DeviceState.Q_ACTUATED: lambda: d.qpos[self.joint_ids],
            DeviceState.DQ: lambda: d.qvel[self.joint_ids_all],
            DeviceState.DQ_ACTUATED: lambda: d.qvel[self.joint_ids],
            DeviceState.DDQ: lambda: d.qacc[self.joint_ids_all],
            DeviceState.EE_XYZ: lambda: d.get_body_xpos(self.EE),
            DeviceState.EE_XYZ_VEL: lambda: d.get_body_xvelp(self.EE),
            DeviceState.EE_QUAT: lambda: d.get_body_xquat(self.EE),
            DeviceState.FORCE: lambda: self._wrench_part(0),
            DeviceState.TORQUE: lambda: self._wrench_part(1),
            DeviceState.J: lambda: self.jacobian(),
        }
        self._cache: Dict[DeviceState, Any] = {}
        self._locks: Dict[DeviceState, Lock] = {key: Lock() for key in DeviceState}
        self.concise_state_vars = [
            DeviceState.Q_ACTUATED, DeviceState.DQ_ACTUATED, DeviceState.EE_XYZ,
            DeviceState.EE_XYZ_VEL, DeviceState.EE_QUAT, DeviceState.FORCE, DeviceState.TORQUE,
        ]

    # ------------------------------------------------------------------
    def jacobian(self, full: bool = False):
        """[jacp; jacr] of the EE body (6 x nv), rows masked by ctrlr_dof unless `full`
        (device.py:115-133)."""
        d = self.sim.data
        J = np.vstack([np.asarray(d.get_body_jacp(self.EE)).reshape(3, -1),
                       np.asarray(d.get_body_jacr(self.EE)).reshape(3, -1)])
        return J if full else J[self.ctrlr_dof]

    def ft_frame_xmat(self):
        """World orientation of the wrist F/T site, None for devices without one (device.py:135-143)."""
        if self.name in _FT_SLICES:
            return self.sim.data.get_site_xmat("ft_frame_" + self.name)
        return None

    def ft_raw(self):
        """Sensor-frame [force, torque] (6,), zeros for devices without a sensor."""
        if self.name in _FT_SLICES:
            f, t = _FT_SLICES[self.name]
            sd = self.sim.data.sensordata
            return np.concatenate([sd[f], sd[t]])
        return np.zeros(6)

    def _wrench_part(self, which: int):
        if self.name not in _FT_SLICES:
            return np.zeros(3)
        sl = _FT_SLICES[self.name][which]
        return np.matmul(self.ft_frame_xmat(), self.sim.data.sensordata[sl])

    # ------------------------------------------------------------------
    def get_state(self, state_var: DeviceState):
        if self._use_sim:
            return copy.copy(self._getters[state_var]())
        with self._locks[state_var]:
            return copy.copy(self._cache[state_var])

    def _refresh(self, state_var: DeviceState):
        assert self._use_sim is False
        with self._locks[state_var]:
            self._cache[state_var] = copy.copy(self._getters[state_var]())

    def get_all_states(self):
        return {key: self.get_state(key) for key in self.concise_state_vars}

    def update_state(self):
        """Polling-thread body (device.py:199-205): only legal when `use_sim` is False."""
        assert self._use_sim is False
        for var in DeviceState:
            self._refresh(var)

    def get_all_joint_ids(self):
        return self.joint_ids_all

    def get_actuator_joint_ids(self):
        return self.joint_ids

    def get_gripper_joint_ids(self):
        return self.gripper_ids
