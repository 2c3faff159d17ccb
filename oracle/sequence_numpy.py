"""TEST INFRASTRUCTURE ONLY - CPU restatement of the caller loop of the reference's insertion demo.

Follows ir-lab/irl_control `examples/insertion_task.py` statement by statement:
    run_sequence            312-317
    go_to_waypoint          279-297
    grip                    190-204   (the wall-clock timer is counted in control steps here)
    send_forces             146-179   (gripper override, error update)
    set_waypoint_targets    206-268   (passive arm: current EE xyz + DEFAULT_EE_QUAT)
The simulator is replaced by a pose stream: `poses[t]` is the state the t-th `generate` call sees,
`poses[t + 1]` the state after its `sim.step()`.

Pinned against the reference itself: the example as a program cannot run here (mujoco_py, an MjViewer, the
scene meshes), but its methods can - `oracle/ref_harness.drive_reference_sequence` /
`drive_reference_gain_test` / `reference_object_placement` call the UNMODIFIED `run_sequence`, `go_to_waypoint`,
`grip`, `send_forces`, `set_waypoint_targets`, `initialize_action_objects[_random]` and `GainTest.run` on a
subclass whose simulator, viewer and wall-clock timer are stand-ins, and the tests compare this restatement with
them tick by tick (tests/test_insertion_host.py, tests/test_sequence_host.py; build container only).
"""
import numpy as np

from . import osc_numpy


class _Stop(Exception):
    pass


def run_sequence(actions, wp_xyz, wp_quat, poses, active_dev, passive_quat, max_vel0, n_ticks):
    """actions: list of dicts with the defaults applied (kp, max_error, min_speed_xyz, max_speed_xyz,
    gripper_force, grip_steps); wp_xyz / wp_quat [A, .]: what set_waypoint_targets computes for the
    active arm per action; poses: dict of arrays indexed by tick - active_xyz, active_quat,
    passive_xyz; active_dev: device dict of the layout (for calc_error).
    Returns one record per generate() call."""
    rec = []
    tick = [0]
    targets = {"active_xyz": np.zeros(3), "active_quat": np.array([1.0, 0, 0, 0]),      # Target() (utils.py:10-15)
               "passive_xyz": np.zeros(3), "passive_quat": np.array([1.0, 0, 0, 0])}
    state = {"max_vel0": float(max_vel0), "err": 0.0}

    def calc_error_norm():
        t = tick[0]
        return float(np.linalg.norm(osc_numpy.calc_error(active_dev, poses["active_xyz"][t], poses["active_quat"][t],
                                                         targets["active_xyz"], targets["active_quat"])))

    def generate_and_send(a, gripper_force):
        # controller.generate(self.targets) sees the state of tick t; send_forces steps the simulator
        rec.append(dict(tick=tick[0], action=a, err=state["err"], max_vel0=state["max_vel0"],
                        gripper_force=float(gripper_force) if gripper_force else 0.0,
                        **{k: np.array(v, dtype=np.float64) for k, v in targets.items()}))
        tick[0] += 1
        if tick[0] >= n_ticks:
            raise _Stop
        state["err"] = calc_error_norm()                      # insertion_task.py:169-179

    try:
        for a, p in enumerate(actions):
            if p["action"] == "WP":
                # set_waypoint_targets (206-268)
                targets["passive_xyz"] = np.array(poses["passive_xyz"][tick[0]], dtype=np.float64)
                targets["passive_quat"] = np.array(passive_quat, dtype=np.float64)
                targets["active_xyz"] = np.array(wp_xyz[a], dtype=np.float64)
                targets["active_quat"] = np.array(wp_quat[a], dtype=np.float64)
                state["err"] = np.inf                         # 289
                while state["err"] > p["max_error"]:          # 290
                    state["max_vel0"] = max(p["min_speed_xyz"], min(p["max_speed_xyz"], p["kp"] * state["err"]))
                    generate_and_send(a, p["gripper_force"])
            else:
                for _ in range(p["grip_steps"]):              # while self.timer_running (201)
                    generate_and_send(a, p["gripper_force"])
        while True:                                           # sequence finished: the batch keeps holding
            generate_and_send(len(actions), 0.0)
    except _Stop:
        pass
    return rec


def run_waypoint_cycle(wps, ee_xyz, threshold, n_ticks):
    """examples/gain_test.py:134-162 for ONE arm: `wps [W, 3]` its waypoint list, `ee_xyz[t]` the EE position
    the t-th generate() sees.  Returns (target, wp_idx before the update) per tick."""
    idx, rec = 0, []
    for t in range(n_ticks):
        target = np.array(wps[idx], dtype=np.float64)            # targets[...].set_xyz(wps[idx])       (138-139)
        rec.append((target, idx))                                # controller.generate(targets)         (143)
        err = np.linalg.norm(ee_xyz[t] - target)                 # self.errors[...]                     (151-152)
        if err < threshold:                                      # (153-162)
            idx = idx + 1 if idx < len(wps) - 1 else 0
    return rec


# ---------------------------------------------------------------- around the loop: object placement, waypoint targets
# (restated one episode at a time with the transforms3d restatement of oracle/t3d.py)
DEFAULT_EE_ROT = np.deg2rad([0, -90, -90])                        # insertion_task.py:18


def initialize_action_objects(action_objects):
    """Configured placement (insertion_task.py:299-311): YAML position, YAML Euler angles in degrees -> radians ->
    quaternion.  Returns {object: (pos, quat)} - what `set_free_joint_qpos` would store."""
    from . import t3d
    placed = {}
    for name, spec in action_objects.items():
        angles = np.deg2rad(spec['initial_pos_abg'])
        placed[name] = (np.array(spec['initial_pos_xyz'], dtype=np.float64), np.array(t3d.euler2quat(*angles)))
    return placed


def initialize_action_objects_random(action_objects, arm_name, uniform):
    """Randomised placement (insertion_task.py:341-369).  `uniform(low, high)` stands for `np.random.uniform`; the
    reference draws, in this order: male x, male y, female x, female y, male yaw, female yaw.  x is mirrored for the
    left arm, z stays as configured, the yaw is cut to an integer and - unlike the configured placement - handed to
    euler2quat WITHOUT a degree conversion (358, 367).  `action_objects` is updated in place like the reference's."""
    from . import t3d
    xy = {"male_object": (uniform(0.4, 0.6), uniform(0.5, 0.7))}
    xy["female_object"] = (uniform(0.0, 0.3), uniform(0.5, 0.7))
    side = 1.0 if arm_name == 'right' else -1.0
    placed = {}
    for name in ("male_object", "female_object"):
        spec = action_objects[name]
        spec['initial_pos_abg'] = [0, 0, int(uniform(-20, 20))]
        spec['initial_pos_xyz'][0] = side * xy[name][0]
        spec['initial_pos_xyz'][1] = xy[name][1]
        placed[name] = (np.array(spec['initial_pos_xyz'], dtype=np.float64),
                        np.array(t3d.euler2quat(*spec['initial_pos_abg'])))
    return placed


def set_waypoint_targets(params, action_objects, object_qpos, start_pos):
    """Active-arm target of one WP action (insertion_task.py:217-270).  `object_qpos[name] = (pos, quat)`: the free
    joint's qpos as the simulator reports it when the waypoint starts.  Returns (xyz, quat).

    Position: `start_pos`, or an object's position plus an offset that is either a 3-list or the NAME of one of that
    object's offset attributes; a literal list target is concatenated with the offset list by the reference's `+`
    and then rejected by `Target.set_xyz` (len != 3).  Orientation: none -> DEFAULT_EE_QUAT; a list -> degrees;
    an object name -> R(object) . R(DEFAULT_EE_ROT + [0, 0, grip_yaw]) -> static-xyz Euler angles -> quaternion."""
    from . import t3d
    if 'target_xyz' not in params:
        raise KeyError('target_xyz')
    where, shift = params['target_xyz'], params.get('offset', [0.0, 0.0, 0.0])
    if isinstance(where, str):
        if where == 'start_pos':
            xyz = np.asarray(start_pos, dtype=np.float64)
        else:
            if isinstance(shift, str):
                shift = action_objects[where][shift]
            xyz = np.asarray(object_qpos[where][0], dtype=np.float64) + np.asarray(shift, dtype=np.float64)
    elif isinstance(where, list):
        joined = where + shift                                     # list concatenation, as in the reference
        assert len(joined) == 3                                    # Target.set_xyz (utils.py:36)
        xyz = np.asarray(joined, dtype=np.float64)
    else:
        raise ValueError
    facing = params.get('target_abg')
    if facing is None:
        return xyz, np.asarray(t3d.euler2quat(*DEFAULT_EE_ROT))
    if isinstance(facing, str):
        yaw = np.deg2rad(action_objects[facing]['grip_yaw'])
        rot = t3d.quat2mat(object_qpos[facing][1]) @ t3d.euler2mat(*(DEFAULT_EE_ROT + [0, 0, yaw]))
        abg = np.array(t3d.mat2euler(rot))
    elif isinstance(facing, list):
        abg = np.deg2rad(facing)
    else:
        raise ValueError
    return xyz, np.asarray(t3d.euler2quat(*abg))                  # Target.set_abg (utils.py:52-54)
