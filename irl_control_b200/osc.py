"""`OSC` - operational-space controller with the reference's Python API, computed on the GPU.

Drop-in for `irl_control.OSC` (irl_control/osc.py:13-210): same constructor
arguments, `generate(targets) -> (force_idxs, forces)` with lists ordered like
`targets`, public `calc_error(target, device)`.  The control law itself is
NOT evaluated here: `generate` gathers one robot's state exactly as
`Robot.get_all_states` does, hands it to the CUDA library as a batch of one,
and unpacks the result.  `generate_batch` is the entry point the B200 path
exists for: B independent robot instances per call, state resident in HBM.

Differences from the reference that a caller can observe: none intended.  In particular the gain vectors
`task_space_gains` / `lamb` are the ones `__init__` stored in the (shared, mutable) config dict (osc.py:35-39): like
the reference, a later edit of cfg['kp'] / ['kv'] / ['ko'] changes the saturation and the kv factors but not them.
Quirks kept on purpose: per-device overwrite of the
velocity term, `np.all(target_vel) == 0` branch rule, J_idxs in sub-device
order (an out-of-range index there raises IndexError like numpy would).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import _native
from .device import Device, DeviceState
from .engine import BatchedOSC
from .layout import OscLayout, compile_layout
from .robot import Robot, RobotState
from .rotations import normalized_vector, qconjugate, qmult, quat2euler
from .utils import ControllerConfig, Target


class OSC:
    def __init__(self, robot: Robot, sim, input_device_configs: Sequence[Tuple[str, Dict]],
                 nullspace_config: Optional[Dict] = None, use_g=True, admittance=False):
        self.sim = sim
        self.robot = robot
        self.device_configs: Dict[str, ControllerConfig] = {
            name: ControllerConfig(cfg) for name, cfg in input_device_configs}
        self.nullspace_config = nullspace_config
        self.use_g = use_g
        self.admittance = admittance
        # osc.py:35-39 - the derived gain vectors are written into the caller's dicts
        for cfg in self.device_configs.values():
            kv, kp, ko = cfg.get_params(['kv', 'kp', 'ko'])
            gains = np.array([kp] * 3 + [ko] * 3)
            cfg['task_space_gains'] = gains
            cfg['lamb'] = gains / kv
        self._engines: Dict[tuple, BatchedOSC] = {}

    # ------------------------------------------------------------------
    def _signature(self, names: Sequence[str]) -> tuple:
        """Everything a compiled layout depends on that callers may mutate between steps."""
        sig = [tuple(names)]
        for name in names:
            dev = self.robot.sub_devices_dict[name]
            cfg = self.device_configs[name]
            sig.append((dev.max_vel is None, cfg['kp'], cfg['kv'], cfg['ko'], tuple(cfg['k']), tuple(cfg['d']),
                        tuple(float(x) for x in cfg['task_space_gains']), tuple(float(x) for x in cfg['lamb'])))
        ns = None if self.nullspace_config is None else self.nullspace_config['kv']
        sig.append((ns, bool(self.use_g), self.admittance is True))
        return tuple(sig)

    def layout_for(self, target_names: Sequence[str]) -> OscLayout:
        cfgs = {name: self.device_configs[name].ctrlr_dict for name in target_names}
        return compile_layout(self.robot, cfgs, target_names, self.nullspace_config,
                              bool(self.use_g), self.admittance is True)

    def engine_for(self, target_names: Sequence[str]) -> BatchedOSC:
        key = self._signature(target_names)
        eng = self._engines.get(key)
        if eng is None:
            eng = BatchedOSC(self.layout_for(target_names))
            self._engines[key] = eng
        return eng

    # ------------------------------------------------------------------
    def calc_error(self, target: Target, device: Device):
        """Pose error of one device (osc.py:101-118): [ee - target ; euler of the quaternion error].

        Host-side scalar version for callers that use it as a stop criterion
        (insertion_task.py:173-179); the batched device version is `BatchedOSC.calc_error`.
        """
        u_task = np.zeros(6)
        if np.sum(device.ctrlr_dof_xyz) > 0:
            u_task[:3] = device.get_state(DeviceState.EE_XYZ) - target.get_xyz()
        if np.sum(device.ctrlr_dof_abg) > 0:
            q_d = normalized_vector(target.get_quat())
            q_r = qmult(q_d, qconjugate(device.get_state(DeviceState.EE_QUAT)))
            u_task[3:] = quat2euler(qconjugate(q_r))
        return u_task

    # ------------------------------------------------------------------
    def gather_state(self, targets: Dict[str, Target]) -> Dict[str, np.ndarray]:
        """One robot's inputs as a batch of one, fields in target order.  Pulls exactly the state variables the control
        law reads (robot.py:125-136 pulls all of them: M, DQ, J, and per target device EE_XYZ, EE_QUAT [+ FORCE, TORQUE])
        through the same `get_state` accessors, i.e. from the simulator or, in polling-thread mode, from the cache."""
        names = list(targets.keys())
        robot = self.robot
        Js, _ = robot.get_state(RobotState.J)
        D = len(names)
        st = {
            "M": np.ascontiguousarray(robot.get_state(RobotState.M), dtype=np.float64)[None],
            "J": np.ascontiguousarray(np.vstack([Js[nm] for nm in names]), dtype=np.float64)[None],
            "dq": np.ascontiguousarray(robot.get_state(RobotState.DQ), dtype=np.float64)[None],
            "ee_xyz": np.zeros((1, D, 3)), "ee_quat": np.zeros((1, D, 4)),
            "target_xyz": np.zeros((1, D, 3)), "target_quat": np.zeros((1, D, 4)),
            "max_vel": np.zeros((1, D, 2)),
        }
        if self.use_g:
            st["bias"] = np.ascontiguousarray(
                np.asarray(self.sim.data.qfrc_bias)[self.robot.joint_ids_all], dtype=np.float64)[None]
        tvel = np.zeros((1, D, 6))
        if self.admittance is True:
            st["ft_xmat"] = np.zeros((1, D, 9))
            st["ft_raw"] = np.zeros((1, D, 6))
        for d, nm in enumerate(names):
            dev = self.robot.get_device(nm)
            tgt = targets[nm]
            st["ee_xyz"][0, d] = dev.get_state(DeviceState.EE_XYZ)
            st["ee_quat"][0, d] = dev.get_state(DeviceState.EE_QUAT)
            st["target_xyz"][0, d] = tgt.get_xyz()
            st["target_quat"][0, d] = tgt.get_quat()
            tvel[0, d] = np.hstack([tgt.get_xyz_vel(), tgt.get_abg_vel()])
            if dev.max_vel is not None:
                st["max_vel"][0, d] = dev.max_vel
            if self.admittance is True:
                if self.robot.is_using_sim():
                    R = dev.ft_frame_xmat()
                    st["ft_xmat"][0, d] = np.eye(3).reshape(-1) if R is None else np.asarray(R).reshape(-1)
                    st["ft_raw"][0, d] = dev.ft_raw()
                else:
                    # polling-thread mode: the wrench of the SAME snapshot as the rest of the state (osc.py:179 reads
                    # robot_state[name][FORCE / TORQUE]), already rotated into the world frame
                    st["ft_xmat"][0, d] = np.eye(3).reshape(-1)
                    st["ft_raw"][0, d] = np.concatenate([dev.get_state(DeviceState.FORCE), dev.get_state(DeviceState.TORQUE)])
        if np.any(tvel != 0.0):
            st["target_vel"] = tvel
        return st

    def generate(self, targets: Dict[str, Target]):
        """Forces for the devices named in `targets` (osc.py:120-210)."""
        if self.robot.is_using_sim() is False:
            assert self.robot.is_running(), "Robot must be running!"
        names = list(targets.keys())
        engine = self.engine_for(names)
        out = engine.step_host(self.gather_state(targets))
        status = int(out["status"][0])
        if status & _native.ST_DX_RANGE:
            k = engine.k
            raise IndexError("index %d is out of bounds for axis 0 with size %d" % (k, k))
        forces, force_idxs = [], []
        for sl, dl in zip(engine.layout.ctrl_slices, engine.layout.devices):
            forces.append(out["ctrl"][0, sl].copy())
            force_idxs.append(self.robot.sub_devices_dict[dl.name].ctrl_idxs)
        return force_idxs, forces

    def generate_batch(self, target_names: Sequence[str], state: Dict, **kw) -> Dict:
        """B instances at once; `state` holds CUDA tensors (-> BatchedOSC.step) or numpy
        arrays (-> BatchedOSC.step_host), per-device fields in `target_names` order."""
        engine = self.engine_for(list(target_names))
        first = state["M"] if "M" in state else state["qM"]       # qM: MuJoCo's sparse inertia (IRLOSC_M_QM)
        if isinstance(first, np.ndarray):
            return engine.step_host(state, **kw)
        return engine.step(state, **kw)
