"""`MujocoApp` - base class of the demo applications (reference: mujoco_app.py:10-63).

Same constructor and helper methods.  The simulator behind `self.sim` is
pluggable: the reference hard-wires `mujoco_py.MjSim`; here a `sim_factory`
may be injected, and the default is `SyntheticSim` (no MuJoCo in this image).
Scene XML files are not parsed - a scene name only selects how many free
bodies follow the robot (`configs.SCENE_FREE_OBJECTS`).
"""
import os
import time
from typing import Callable, Dict, Optional

import numpy as np
import yaml

from . import configs
from .device import Device
from .dual_ur5 import DualUR5Model
from .robot import Robot
from .sim import SyntheticSim


class MujocoApp:
    def __init__(self, robot_config_file: str = None, scene_file: str = None, use_sim: bool = True,
                 sim_factory: Optional[Callable[[str], object]] = None, config_override: Optional[Dict] = None):
        # config_override: an already loaded config dict (same schema) instead of a file / built-in name
        self.config = config_override if config_override is not None else self._load_config(robot_config_file)
        if sim_factory is not None:
            self.sim = sim_factory(scene_file)
        else:
            n_free = configs.SCENE_FREE_OBJECTS.get(os.path.basename(scene_file or ""), 0)
            self.sim = SyntheticSim(DualUR5Model(n_free_objects=n_free))
        self.model = self.sim.model
        self.devices = np.array([Device(dev, self.model, self.sim, use_sim) for dev in self.config['devices']])
        self.create_robot_devices(self.config['robots'], use_sim)
        self.controller_configs = self.config['controller_configs']
        self.timer_running = False

    @staticmethod
    def _load_config(robot_config_file: str) -> Dict:
        if robot_config_file is not None and os.path.isfile(robot_config_file):
            with open(robot_config_file, 'r') as fh:
                return yaml.safe_load(fh)
        return configs.robot_config(robot_config_file)

    def create_robot_devices(self, robot_yml, use_sim: bool):
        """Replace the devices listed under each robot entry by one `Robot` (mujoco_app.py:24-35)."""
        robots, grouped = [], []
        for entry in robot_yml:
            ids = list(entry['device_ids'])
            grouped += ids
            robots.append(Robot(list(self.devices[ids]), entry['name'], self.sim, use_sim))
        loose = [self.devices[i] for i in range(len(self.devices)) if i not in set(grouped)]
        self.devices = np.array(loose + robots, dtype=object)

    def sleep_for(self, sleep_time: float):
        """Blocking timer other threads poll through `timer_running` (mujoco_app.py:37-41); not re-entrant."""
        assert self.timer_running == False  # noqa: E712 (same contract as the reference)
        self.timer_running = True
        try:
            time.sleep(sleep_time)
        finally:
            self.timer_running = False

    def get_robot(self, robot_name: str) -> Robot:
        """The `Robot` with that name, or None (mujoco_app.py:43-47)."""
        return next((item for item in self.devices if type(item) == Robot and item.name == robot_name), None)

    def get_controller_config(self, name: str) -> Dict:
        """The `controller_configs` entry with that name - the YAML dict itself, not a copy - or None."""
        return next((entry for entry in self.config['controller_configs'] if entry['name'] == name), None)

    def set_free_joint_qpos(self, free_joint_name, quat=None, pos=None):
        """Write a free joint's pose into `sim.data.qpos` (layout: x y z, then w x y z; mujoco_app.py:55-63)."""
        start = self.sim.model.jnt_qposadr[self.sim.model.joint_name2id(free_joint_name)]
        for value, lo, hi in ((pos, 0, 3), (quat, 3, 7)):
            if value is not None:
                self.sim.data.qpos[start + lo:start + hi] = value
