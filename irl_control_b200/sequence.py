"""Batched action sequences - the caller loop of the reference's insertion demo on the GPU.

Reference: `examples/insertion_task.py` - `run_sequence` (312-317) walks a list of WP / GRIP actions
(`action_sequence_configs/insertion_task.yaml:35-104`); `go_to_waypoint` (279-297) loops
`generate -> send_forces` until the active arm's pose error is <= `max_error`, setting
`active_arm.max_vel[0] = clip(kp * error, min_speed_xyz, max_speed_xyz)` every step; `grip`
(190-204) holds for `gripper_duration`; `send_forces` (146-179) overrides the gripper's ctrl slot
and updates the error; `set_waypoint_targets` (206-268) fixes the passive arm at the xyz it has
when a waypoint starts.  Here one `BatchedOSC.step_sequence` call advances B independent episodes
by one control step: the state machine runs inside the fused step kernel (`osc_sequence.cuh`).

What stays with the caller (it needs the simulator's objects): the waypoint poses themselves -
`wp_xyz / wp_quat [B, n_actions, .]`, what `set_waypoint_targets` computes from the object poses
and the YAML offsets.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from . import _native

# examples/insertion_task.py:82-103 get_default_action_ctrl_params
WP_DEFAULTS = {"kp": 6.0, "max_error": 0.0018, "gripper_force": 0.0, "min_speed_xyz": 0.1, "max_speed_xyz": 3.0}
# the reference spells the second key 'gripper_duation' (insertion_task.py:101), so a GRIP entry without its own
# `gripper_duration` raises KeyError there (196) - and here
GRIP_DEFAULTS = {"gripper_force": -0.08, "gripper_duation": 1.0}
# examples/insertion_task.py:18-20: euler2quat(*deg2rad([0, -90, -90])), static xyz
DEFAULT_EE_ROT = np.deg2rad([0.0, -90.0, -90.0])


def default_ee_quat() -> np.ndarray:
    from .rotations import euler2quat
    return np.asarray(euler2quat(*DEFAULT_EE_ROT), dtype=np.float64)


class ActionSequence:
    """Compiled action list for one (layout, active arm)."""

    def __init__(self, layout, actions: Sequence[Dict], active_arm: str, step_period: float = 0.002):
        names = [d.name for d in layout.devices]
        if active_arm not in names:
            raise KeyError(active_arm)
        self.layout = layout
        self.active_device = names.index(active_arm)
        self.n_actions = len(actions)
        if not 1 <= self.n_actions <= _native.MAX_ACTIONS:
            raise ValueError("1..%d actions" % _native.MAX_ACTIONS)
        # send_forces: gripper_idx = 7 (ur5right) / 14 (ur5left) in sim.data.ctrl (insertion_task.py:152-156)
        gripper_ctrl = {"ur5right": 7, "ur5left": 14}[active_arm]
        dl = layout.devices[self.active_device]
        self.gripper_slot = -1
        if gripper_ctrl in dl.ctrl_idxs:
            self.gripper_slot = layout.ctrl_slices[self.active_device].start + list(dl.ctrl_idxs).index(gripper_ctrl)
        self.params: List[Dict] = []
        c = _native.Sequence()
        c.n_actions, c.active_device, c.gripper_slot = self.n_actions, self.active_device, self.gripper_slot
        for i, v in enumerate(default_ee_quat()):
            c.passive_quat[i] = float(v)
        for a, entry in enumerate(actions):
            kind = entry["action"]
            p = dict(entry)
            if kind == "WP":
                for k_, v in WP_DEFAULTS.items():                 # update_action_ctrl_params (272-277)
                    p.setdefault(k_, v)
                c.action[a].type = _native.ACT_WP
                c.action[a].kp, c.action[a].max_error = float(p["kp"]), float(p["max_error"])
                c.action[a].min_speed_xyz, c.action[a].max_speed_xyz = float(p["min_speed_xyz"]), float(p["max_speed_xyz"])
            elif kind == "GRIP":
                for k_, v in GRIP_DEFAULTS.items():
                    p.setdefault(k_, v)
                c.action[a].type = _native.ACT_GRIP
                # the reference sleeps `gripper_duration` seconds of wall clock while stepping; in a batch the
                # duration is counted in control steps
                p["grip_steps"] = int(round(float(p["gripper_duration"]) / step_period))
                c.action[a].grip_steps = p["grip_steps"]
            else:
                raise ValueError("unknown action %r" % kind)
            c.action[a].gripper_force = float(p["gripper_force"])
            self.params.append(p)
        self.c_struct = c

    def new_state(self, B: int, wp_xyz, wp_quat, device=None) -> Dict:
        """Episode state for B instances.  `wp_xyz [B, A, 3]`, `wp_quat [B, A, 4]`: active-arm targets per
        action (rows of GRIP actions are ignored).  torch tensors on `device`, or numpy when device is None."""
        D = self.layout.D
        mv0 = float(self.layout.devices[self.active_device].max_vel[0])
        if device is None:
            st = {"action": np.zeros(B, np.int32), "entered": np.zeros(B, np.int32), "timer": np.zeros(B, np.int32),
                  "err": np.zeros(B), "max_vel0": np.full(B, mv0), "target_xyz": np.zeros((B, D, 3)),
                  "target_quat": np.zeros((B, D, 4))}
            st["target_quat"][..., 0] = 1.0                   # Target(): quat [1, 0, 0, 0] (utils.py:10-15)
            st["wp_xyz"] = np.ascontiguousarray(wp_xyz, dtype=np.float64)
            st["wp_quat"] = np.ascontiguousarray(wp_quat, dtype=np.float64)
        else:
            import torch
            i32 = dict(dtype=torch.int32, device=device)
            f64 = dict(dtype=torch.float64, device=device)
            st = {"action": torch.zeros(B, **i32), "entered": torch.zeros(B, **i32), "timer": torch.zeros(B, **i32),
                  "err": torch.zeros(B, **f64), "max_vel0": torch.full((B,), mv0, **f64),
                  "target_xyz": torch.zeros(B, D, 3, **f64), "target_quat": torch.zeros(B, D, 4, **f64)}
            st["target_quat"][..., 0] = 1.0
            st["wp_xyz"] = torch.as_tensor(wp_xyz, **f64).contiguous()
            st["wp_quat"] = torch.as_tensor(wp_quat, **f64).contiguous()
        assert tuple(st["wp_xyz"].shape) == (B, self.n_actions, 3) and tuple(st["wp_quat"].shape) == (B, self.n_actions, 4)
        return st
