"""`Robot` - a group of `Device`s sharing one joint space (host side).

Mirrors the reference's `irl_control/robot.py:16-144`: same constructor,
`joint_ids_all` / `num_joints_total`, `get_state(RobotState.*)`,
`get_all_states`, the optional 1 kHz polling thread (`start`/`stop`) with its
asserts.  `RobotState.J` keeps the reference's return shape: `(Js, J_idxs)`
with `J_idxs` numbered in SUB-DEVICE order (robot.py:52-55) - the latent order
mismatch against the target-ordered `dx` is preserved and is
what `layout.py` encodes as `dx_idx`.
"""
import copy
import time
from enum import Enum
from threading import Lock
from typing import Any, Callable, Dict, List

import numpy as np

from .device import Device, DeviceState


class RobotState(Enum):
    M = 'INERTIA'
    DQ = 'DQ'
    J = 'JACOBIAN'


def dense_mass_matrix(sim) -> np.ndarray:
    """nv x nv joint-space inertia of the whole scene (what `_mj_fullM` yields, robot.py:69-70).

    A backend may expose `sim.full_mass_matrix()` (SyntheticSim); with real
    mujoco_py the sparse `qM` is expanded by `cymj._mj_fullM`.
    """
    if hasattr(sim, "full_mass_matrix"):
        return sim.full_mass_matrix()
    import mujoco_py as mjp  # only reached with a real simulator
    nv = sim.model.nv
    buf = np.zeros(nv * nv)
    mjp.cymj._mj_fullM(sim.model, buf, sim.data.qM)
    return buf.reshape(nv, nv)


class Robot:
    def __init__(self, sub_devices: List[Device], robot_name, sim, use_sim, collect_hz=1000):
        self.sim = sim
        self._use_sim = use_sim
        self.sub_devices = sub_devices
        self.sub_devices_dict: Dict[str, Device] = {dev.name: dev for dev in sub_devices}
        self.name = robot_name
        self.num_scene_joints = self.sim.model.nv
        ids = np.array([], dtype=np.int32)
        for dev in self.sub_devices:
            ids = np.hstack([ids, dev.joint_ids_all])
        self.joint_ids_all = np.sort(np.unique(ids))
        self.num_joints_total = len(self.joint_ids_all)
        self.running = False
        self.data_collect_hz = collect_hz
        self._getters: Dict[RobotState, Callable[[], Any]] = {
            RobotState.M: self.mass_matrix,
            RobotState.DQ: self.joint_velocities,
            RobotState.J: self.jacobians,
        }
        self._cache: Dict[RobotState, Any] = {}
        self._locks: Dict[RobotState, Lock] = {key: Lock() for key in RobotState}

    # ---- state pulls (robot.py:44-72) ----------------------------------
    def jacobians(self):
        Js, J_idxs = {}, {}
        row = 0
        for name, device in self.sub_devices_dict.items():
            Jd = device.get_state(DeviceState.J)
            J_idxs[name] = np.arange(row, row + Jd.shape[0])
            row += Jd.shape[0]
            Js[name] = Jd[:, self.joint_ids_all]
        return Js, J_idxs

    def joint_velocities(self):
        dq = np.zeros(self.joint_ids_all.shape)
        for dev in self.sub_devices:
            # global joint ids used as local positions, as in the reference (robot.py:64)
            dq[dev.get_all_joint_ids()] = dev.get_state(DeviceState.DQ)
        return dq

    def mass_matrix(self):
        M = dense_mass_matrix(self.sim)
        return M[np.ix_(self.joint_ids_all, self.joint_ids_all)]

    # ---- cached / live access ------------------------------------------
    def get_state(self, state_var: RobotState):
        if self._use_sim:
            return copy.copy(self._getters[state_var]())
        with self._locks[state_var]:
            return copy.copy(self._cache[state_var])

    def _refresh(self, state_var: RobotState):
        assert self._use_sim is False
        with self._locks[state_var]:
            self._cache[state_var] = copy.copy(self._getters[state_var]())

    def is_running(self):
        return self.running

    def is_using_sim(self):
        return self._use_sim

    def start(self):
        """Polling loop (robot.py:103-116); intended as a thread target."""
        assert self.running is False and self._use_sim is False
        self.running = True
        period = 1.0 / float(self.data_collect_hz)
        last = time.time()
        while self.running:
            for dev in self.sub_devices:
                dev.update_state()
            for var in RobotState:
                self._refresh(var)
            now = time.time()
            time.sleep(max(period - (now - last), 0))
            last = now

    def stop(self):
        assert self.running is True and self._use_sim is False
        self.running = False

    def get_device(self, device_name: str) -> Device:
        return self.sub_devices_dict[device_name]

    def get_device_states(self):
        return {name: dev.get_all_states() for name, dev in self.sub_devices_dict.items()}

    def get_all_states(self):
        state = self.get_device_states()
        for key in RobotState:
            state[key] = self.get_state(key)
        return state
