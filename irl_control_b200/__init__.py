"""irl_control_b200 - B200-native batched operational-space control for the DualUR5.

Public surface mirrors `irl_control/__init__.py:2-5` (Device, Robot, OSC,
MujocoApp) and adds the batched engine.  The CUDA library
(`libirlosc.so`, C ABI in include/irlosc.h) is the only compute path.
"""
from .version import __version__
from .device import Device, DeviceState
from .robot import Robot, RobotState
from .utils import Target, ControllerConfig
from .osc import OSC
from .mujoco_app import MujocoApp
from .engine import BatchedOSC
from .layout import OscLayout, DeviceLayout, compile_layout

__all__ = ["Device", "DeviceState", "Robot", "RobotState", "Target", "ControllerConfig", "OSC",
           "MujocoApp", "BatchedOSC", "OscLayout", "DeviceLayout", "compile_layout", "__version__"]
