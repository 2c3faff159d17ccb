"""Host side of the batched insertion demo: what the reference's `InsertionTask` does AROUND the control
loop, for B episodes at once (reference: `examples/insertion_task.py`, `action_sequence_configs/insertion_task.yaml`).

    object placement      `initialize_action_objects` (299-311), `initialize_action_objects_random` (341-369)
    waypoint poses        `set_waypoint_targets` (206-268), active arm: xyz = object position + offset,
                          orientation = object rotation x (DEFAULT_EE_ROT + grip yaw) -> Euler -> `Target.set_abg`

The result - `wp_xyz [B, A, 3]`, `wp_quat [B, A, 4]` - is what `ActionSequence.new_state` takes; the control
loop itself (`go_to_waypoint`, `grip`, `send_forces`) then runs inside the fused step kernel
(`BatchedOSC.step_sequence`).  All of this is set-up work done once per batch of episodes, in numpy.

Reference behaviour kept on purpose:
  * the randomised placement passes the yaw in DEGREES straight to `euler2quat` (367, 358: `euler2quat(*[0, 0,
    yaw])` with `yaw = int(uniform(-20, 20))`), while the configured placement converts with `deg2rad` (308);
  * `int()` truncates the yaw toward zero;
  * object poses are read when a waypoint starts (236, 254); with no contact simulation here the objects stay
    where they were placed, so `waypoint_poses` evaluates every action on the placed poses.  A caller with a
    real simulator refreshes `state["wp_xyz"] / ["wp_quat"]` rows of later actions from the live object poses.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from .sequence import DEFAULT_EE_ROT

_EPS4 = 4.0 * float(np.finfo(np.float64).eps)
_EPS = float(np.finfo(np.float64).eps)


# ---------------------------------------------------------------- batched 'sxyz' rotations (w x y z)
def euler2quat_b(e: np.ndarray) -> np.ndarray:
    h = 0.5 * np.asarray(e, dtype=np.float64)
    ci, cj, ck = np.cos(h[..., 0]), np.cos(h[..., 1]), np.cos(h[..., 2])
    si, sj, sk = np.sin(h[..., 0]), np.sin(h[..., 1]), np.sin(h[..., 2])
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    return np.stack([cj * cc + sj * ss, cj * sc - sj * cs, cj * ss + sj * cc, cj * cs - sj * sc], -1)


def euler2mat_b(e: np.ndarray) -> np.ndarray:
    e = np.asarray(e, dtype=np.float64)
    ci, cj, ck = np.cos(e[..., 0]), np.cos(e[..., 1]), np.cos(e[..., 2])
    si, sj, sk = np.sin(e[..., 0]), np.sin(e[..., 1]), np.sin(e[..., 2])
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    m = np.empty(e.shape[:-1] + (3, 3))
    m[..., 0, 0], m[..., 0, 1], m[..., 0, 2] = cj * ck, sj * sc - cs, sj * cc + ss
    m[..., 1, 0], m[..., 1, 1], m[..., 1, 2] = cj * sk, sj * ss + cc, sj * cs - sc
    m[..., 2, 0], m[..., 2, 1], m[..., 2, 2] = -sj, cj * si, cj * ci
    return m


def quat2mat_b(q: np.ndarray) -> np.ndarray:
    q = np.asarray(q, dtype=np.float64)
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    nq = w * w + x * x + y * y + z * z
    tiny = nq < _EPS
    s = 2.0 / np.where(tiny, 1.0, nq)
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz = w * xs, w * ys, w * zs
    xx, xy, xz = x * xs, x * ys, x * zs
    yy, yz, zz = y * ys, y * zs, z * zs
    m = np.empty(q.shape[:-1] + (3, 3))
    m[..., 0, 0], m[..., 0, 1], m[..., 0, 2] = 1.0 - (yy + zz), xy - wz, xz + wy
    m[..., 1, 0], m[..., 1, 1], m[..., 1, 2] = xy + wz, 1.0 - (xx + zz), yz - wx
    m[..., 2, 0], m[..., 2, 1], m[..., 2, 2] = xz - wy, yz + wx, 1.0 - (xx + yy)
    m[tiny] = np.eye(3)
    return m


def mat2euler_b(m: np.ndarray) -> np.ndarray:
    m = np.asarray(m, dtype=np.float64)
    cy = np.sqrt(m[..., 0, 0] ** 2 + m[..., 1, 0] ** 2)
    reg = cy > _EPS4
    ax = np.where(reg, np.arctan2(m[..., 2, 1], m[..., 2, 2]), np.arctan2(-m[..., 1, 2], m[..., 1, 1]))
    ay = np.arctan2(-m[..., 2, 0], cy)
    az = np.where(reg, np.arctan2(m[..., 1, 0], m[..., 0, 0]), 0.0)
    return np.stack([ax, ay, az], -1)


# ---------------------------------------------------------------- object placement
ObjectPoses = Dict[str, Tuple[np.ndarray, np.ndarray]]      # object name -> (xyz [B, 3], quat [B, 4])


def configured_object_poses(B: int, action_objects: Dict) -> ObjectPoses:
    """`initialize_action_objects` (299-311): `initial_pos_xyz`, `initial_pos_abg` in degrees -> `deg2rad`."""
    out = {}
    for name, obj in action_objects.items():
        xyz = np.tile(np.asarray(obj.get("initial_pos_xyz", [0.0, 0.0, 0.0]), dtype=np.float64), (B, 1))
        abg = np.deg2rad(np.asarray(obj.get("initial_pos_abg", [0.0, 0.0, 0.0]), dtype=np.float64))
        out[name] = (xyz, np.tile(euler2quat_b(abg), (B, 1)))
    return out


def random_object_poses(B: int, arm_name: str, action_objects: Dict, rng: Optional[np.random.Generator] = None,
                        u: Optional[np.ndarray] = None) -> ObjectPoses:
    """`initialize_action_objects_random` (341-369) for B episodes.  `arm_name` is 'right' or 'left' as in the
    reference.  `u [B, 6]`: uniforms in [0, 1) standing for the reference's six draws, in its order (male x,
    male y, female x, female y, male yaw, female yaw); drawn from `rng` when not given."""
    if u is None:
        rng = np.random.default_rng() if rng is None else rng
        u = rng.random((B, 6))
    u = np.asarray(u, dtype=np.float64)
    assert u.shape == (B, 6)
    lo = np.array([0.4, 0.5, 0.0, 0.5, -20.0, -20.0])
    hi = np.array([0.6, 0.7, 0.3, 0.7, 20.0, 20.0])
    draw = lo + (hi - lo) * u                         # numpy's uniform(low, high) = low + (high - low) * random()
    sign = 1.0 if arm_name == "right" else -1.0
    out = {}
    for name, cx, cy, cyaw in (("male_object", 0, 1, 4), ("female_object", 2, 3, 5)):
        base = np.asarray(action_objects[name].get("initial_pos_xyz", [0.0, 0.0, 0.0]), dtype=np.float64)
        xyz = np.tile(base, (B, 1))
        xyz[:, 0] = sign * draw[:, cx]
        xyz[:, 1] = draw[:, cy]
        yaw = np.trunc(draw[:, cyaw])                 # int(): toward zero
        abg = np.zeros((B, 3))
        abg[:, 2] = yaw                               # degrees handed to euler2quat unchanged (358, 367)
        out[name] = (xyz, euler2quat_b(abg))
    return out


# ---------------------------------------------------------------- waypoint poses
def waypoint_poses(actions: Sequence[Dict], action_objects: Dict, objects: ObjectPoses,
                   start_pos: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Active-arm target per action, `set_waypoint_targets` (206-268) for B episodes.

    actions        the action list of the YAML (`insertion_action_sequence`)
    action_objects its object block (`nist_action_objects` / `grommet_action_objects`)
    objects        placed object poses, see `configured_object_poses` / `random_object_poses`
    start_pos      [B, 3] EE xyz of the active arm when the sequence starts (`run_sequence`, 313)
    Rows of GRIP actions are left at zero / identity (never read)."""
    start_pos = np.asarray(start_pos, dtype=np.float64)
    B, A = start_pos.shape[0], len(actions)
    wp_xyz = np.zeros((B, A, 3))
    wp_quat = np.zeros((B, A, 4))
    wp_quat[..., 0] = 1.0
    default_quat = euler2quat_b(DEFAULT_EE_ROT)
    for a, p in enumerate(actions):
        if p["action"] != "WP":
            continue
        if "target_xyz" not in p:
            raise KeyError("target_xyz")                                        # 245-247
        offset = p.get("offset", [0.0, 0.0, 0.0])                               # 222
        txyz = p["target_xyz"]
        if isinstance(txyz, str):
            if txyz == "start_pos":                                             # 230-231
                wp_xyz[:, a] = start_pos
            else:
                obj = action_objects[txyz]                                      # 234
                if isinstance(offset, str):
                    offset = obj[offset]                                        # 235-236
                wp_xyz[:, a] = objects[txyz][0] + np.asarray(offset, dtype=np.float64)   # 238-239
        elif isinstance(txyz, list):
            # 241: `params['target_xyz'] + offset` concatenates two Python lists into six numbers, which
            # `Target.set_xyz` rejects (`assert len(xyz) == 3`, utils.py:36) - same failure here
            raise AssertionError("list target_xyz + list offset has %d entries" % (len(txyz) + len(offset)))
        else:
            raise ValueError("Invalid type for target_xyz!")                    # 243-244
        if "target_abg" in p:
            tabg = p["target_abg"]
            if isinstance(tabg, str):                                           # 251-263
                obj = action_objects[tabg]
                grip_eul = DEFAULT_EE_ROT + np.array([0.0, 0.0, np.deg2rad(obj["grip_yaw"])])
                tf = quat2mat_b(objects[tabg][1]) @ euler2mat_b(grip_eul)
                abg = mat2euler_b(tf)
            elif isinstance(tabg, list):
                abg = np.tile(np.deg2rad(np.asarray(tabg, dtype=np.float64)), (B, 1))   # 264-265
            else:
                raise ValueError("Invalid type for target_abg!")
            wp_quat[:, a] = euler2quat_b(abg)                                   # Target.set_abg (utils.py:52-54)
        else:
            wp_quat[:, a] = default_quat                                        # 270
    return wp_xyz, wp_quat
