"""`SyntheticSim` - a `mujoco_py.MjSim`-shaped stand-in built on `dual_ur5.dynamics`.

MuJoCo is not installable in this image, and simulating contact physics is
out of scope.  What callers of the reference
API still need from `sim` is served here for ONE robot instance:

    sim.model.*            index tables (`DualUR5Model`)
    sim.data.qpos/qvel/qacc/qfrc_bias/sensordata/ctrl
    sim.data.get_body_xpos/xquat/xvelp/jacp/jacr(name), get_site_xmat(name)
    sim.forward()          recompute everything from (qpos, qvel)
    sim.step()             unconstrained joint-space forward dynamics
                           (semi-implicit Euler, motor actuators only) - enough
                           to close the loop in tests, NOT a MuJoCo replacement
    sim.full_mass_matrix() what `_mj_fullM` would return (nv x nv)
"""
from __future__ import annotations

import numpy as np
import torch

from .dual_ur5 import DualUR5Model, dynamics


class _SimData:
    def __init__(self, model: DualUR5Model):
        self.qpos = np.zeros(model.nq)
        for i in range(model.n_free_objects):
            self.qpos[model.nv_robot + 7 * i + 3] = 1.0
        self.qvel = np.zeros(model.nv)
        self.qacc = np.zeros(model.nv)
        self.qfrc_bias = np.zeros(model.nv)
        self.sensordata = np.zeros(model.nsensordata)
        self.ctrl = np.zeros(model.nu)
        self.qM = None
        self._model = model
        self._dyn = None

    def _body(self, name):
        return self._model.body_name2id(name)

    def get_body_xpos(self, name):
        return self._dyn.xpos[0, self._body(name)].numpy()

    def get_body_xquat(self, name):
        return self._dyn.xquat[0, self._body(name)].numpy()

    def get_body_xvelp(self, name):
        jp, _ = self._dyn.jac_body(self._body(name))
        return jp[0].numpy() @ self.qvel[:self._model.nv_robot]

    def _jac(self, name, which):
        j = self._dyn.jac_body(self._body(name))[which][0].numpy()
        full = np.zeros((3, self._model.nv))
        full[:, :self._model.nv_robot] = j
        return full.reshape(-1)

    def get_body_jacp(self, name):
        return self._jac(name, 0)

    def get_body_jacr(self, name):
        return self._jac(name, 1)

    def get_site_xmat(self, name):
        return self._dyn.site_xmat[0, self._model.site_name2id(name)].numpy()

    def get_joint_qpos(self, name):
        jid = self._model.joint_name2id(name)
        adr = self._model.jnt_qposadr[jid]
        return self.qpos[adr:adr + 7] if jid >= self._model.nv_robot else self.qpos[adr]


class SyntheticSim:
    def __init__(self, model: DualUR5Model = None, timestep: float = 0.001):
        self.model = model if model is not None else DualUR5Model()
        self.data = _SimData(self.model)
        self.timestep = timestep
        self._M = None
        self.forward()

    def forward(self):
        n = self.model.nv_robot
        q = torch.from_numpy(self.data.qpos[:n].copy())[None]
        dq = torch.from_numpy(self.data.qvel[:n].copy())[None]
        dyn = dynamics(self.model, q, dq)
        self.data._dyn = dyn
        M = np.eye(self.model.nv) * 0.1
        M[:n, :n] = dyn.M[0].numpy()
        self._M = M
        self.data.qM = M.reshape(-1)
        self.data.qfrc_bias[:] = 0.0
        self.data.qfrc_bias[:n] = dyn.bias[0].numpy()

    def full_mass_matrix(self) -> np.ndarray:
        return self._M.copy()

    def step(self):
        """q'' = M^-1 (tau - bias) on the robot DoF; free objects stay put."""
        n = self.model.nv_robot
        tau = np.zeros(n)
        for a in range(self.model.nu):
            j = self.model.actuator_trnid[a, 0]
            if self.model.actuator_kind[a] == "position":      # kp = 1 (MuJoCo default)
                tau[j] += self.data.ctrl[a] - self.data.qpos[j]
            else:
                tau[j] += self.data.ctrl[a]
        qacc = np.linalg.solve(self._M[:n, :n], tau - self.data.qfrc_bias[:n])
        self.data.qacc[:n] = qacc
        self.data.qvel[:n] += self.timestep * qacc
        self.data.qpos[:n] += self.timestep * self.data.qvel[:n]
        self.forward()
