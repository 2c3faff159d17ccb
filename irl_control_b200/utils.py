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
    """Set-point of one device: position xyz, orientation as a w-x-y-z quaternion, and their target
    velocities.  Public surface of the reference's class (utils.py:5-67): `get_* / set_*` for `xyz`, `xyz_vel`,
    `quat`, `quat_vel` (stored) and `abg`, `abg_vel` (static-xyz Euler view of the quaternions), plus
    `set_all_quat` / `set_all_abg`.  Setters check the length and keep the caller's array (no copy), like the
    reference's; the accessor methods are generated below from the two tables."""

    _STORED = {"xyz": 3, "xyz_vel": 3, "quat": 4, "quat_vel": 4}
    _EULER_VIEW = {"abg": "quat", "abg_vel": "quat_vel"}

    def __init__(self, xyz_abg=None, xyz_abg_vel=None):
        pose = np.zeros(6) if xyz_abg is None else np.asarray(xyz_abg, dtype=np.float64)
        rate = np.zeros(6) if xyz_abg_vel is None else np.asarray(xyz_abg_vel, dtype=np.float64)
        assert len(pose) == 6 and len(rate) == 6
        self._v = {"xyz": pose[:3].copy(), "xyz_vel": rate[:3].copy(),
                   "quat": euler2quat(*pose[3:]), "quat_vel": euler2quat(*rate[3:])}

    def set_all_quat(self, xyz, quat):
        assert len(xyz) == 3 and len(quat) == 4
        self._v["xyz"], self._v["quat"] = np.asarray(xyz), np.asarray(quat)

    def set_all_abg(self, xyz, abg):
        assert len(xyz) == 3 and len(abg) == 3
        self._v["xyz"], self._v["quat"] = np.asarray(xyz), np.asarray(euler2quat(*abg))

    def velocity6(self) -> np.ndarray:
        """[xyz_vel, abg_vel] exactly as `generate` assembles it (osc.py:172)."""
        return np.hstack([self.get_xyz_vel(), self.get_abg_vel()]).astype(np.float64)


def _stored_accessors(field, size):
    def getter(self):
        return self._v[field]

    def setter(self, value):
        assert len(value) == size
        self._v[field] = np.asarray(value)
    return getter, setter


def _euler_accessors(quat_field):
    def getter(self):
        return np.asarray(quat2euler(self._v[quat_field]))

    def setter(self, angles):
        assert len(angles) == 3
        self._v[quat_field] = np.asarray(euler2quat(*angles))
    return getter, setter


for _field, _size in Target._STORED.items():
    _g, _s = _stored_accessors(_field, _size)
    setattr(Target, "get_" + _field, _g)
    setattr(Target, "set_" + _field, _s)
for _field, _quat_field in Target._EULER_VIEW.items():
    _g, _s = _euler_accessors(_quat_field)
    setattr(Target, "get_" + _field, _g)
    setattr(Target, "set_" + _field, _s)
del _field, _size, _quat_field, _g, _s


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
