"""`Target` and `ControllerConfig` - the input types of `OSC.generate`.

Same public surface as the reference's `irl_control/utils.py:5-80` (method
names, argument meaning, wxyz quaternion storage, 'sxyz' Euler convention) so
caller code written against irl_control runs unchanged; the Euler/quaternion
conversions come from `rotations.py` because transforms3d is not available.
"""
from typing import Any, Dict, Iterable, List

import numpy as np

from .rotations import euler2quat, quat2euler


class Target:
    """Set-point of one device: position xyz, orientation as a w-x-y-z quaternion,
    and their target velocities (utils.py:5-67)."""

    def __init__(self, xyz_abg=None, xyz_abg_vel=None):
        pose = np.zeros(6) if xyz_abg is None else np.asarray(xyz_abg, dtype=np.float64)
        rate = np.zeros(6) if xyz_abg_vel is None else np.asarray(xyz_abg_vel, dtype=np.float64)
        assert len(pose) == 6 and len(rate) == 6
        self._xyz = pose[:3].copy()
        self._xyz_vel = rate[:3].copy()
        self._quat = euler2quat(*pose[3:])
        self._quat_vel = euler2quat(*rate[3:])

    # ---- getters ----
    def get_xyz(self):
        return self._xyz

    def get_xyz_vel(self):
        return self._xyz_vel

    def get_quat(self):
        return self._quat

    def get_quat_vel(self):
        return np.asarray(self._quat_vel)

    def get_abg(self):
        return np.asarray(quat2euler(self._quat))

    def get_abg_vel(self):
        return np.asarray(quat2euler(self._quat_vel))

    # ---- setters ----
    def set_xyz(self, xyz):
        assert len(xyz) == 3
        self._xyz = np.asarray(xyz)

    def set_xyz_vel(self, xyz_vel):
        assert len(xyz_vel) == 3
        self._xyz_vel = np.asarray(xyz_vel)

    def set_quat(self, quat):
        assert len(quat) == 4
        self._quat = np.asarray(quat)

    def set_quat_vel(self, quat_vel):
        assert len(quat_vel) == 4
        self._quat_vel = np.asarray(quat_vel)

    def set_abg(self, abg):
        assert len(abg) == 3
        self._quat = np.asarray(euler2quat(*abg))

    def set_abg_vel(self, abg_vel):
        assert len(abg_vel) == 3
        self._quat_vel = np.asarray(euler2quat(*abg_vel))

    def set_all_quat(self, xyz, quat):
        assert len(xyz) == 3 and len(quat) == 4
        self.set_xyz(xyz)
        self.set_quat(quat)

    def set_all_abg(self, xyz, abg):
        assert len(xyz) == 3 and len(abg) == 3
        self.set_xyz(xyz)
        self.set_abg(abg)

    def velocity6(self) -> np.ndarray:
        """[xyz_vel, abg_vel] exactly as `generate` assembles it (osc.py:172)."""
        return np.hstack([self.get_xyz_vel(), self.get_abg_vel()]).astype(np.float64)


class ControllerConfig:
    """Thin dict wrapper with `get_params` (utils.py:69-80).  It wraps, not copies,
    the dict it is given - `OSC.__init__` writes `task_space_gains` and `lamb`
    into the caller's dict just like the reference does (osc.py:38-39)."""

    def __init__(self, ctrlr_dict: Dict):
        self.ctrlr_dict = ctrlr_dict

    def __getitem__(self, name: str) -> Any:
        return self.ctrlr_dict[name]

    def __setitem__(self, name: str, value: Any) -> None:
        self.ctrlr_dict[name] = value

    def get_params(self, keys: Iterable[str]) -> List[Any]:
        return [self.ctrlr_dict[k] for k in keys]
