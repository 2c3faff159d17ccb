"""Built-in robot / controller configurations of the DualUR5 demos.

The reference keeps these as YAML (`irl_control/robot_configs/*.yaml`); the
same keys and values are expressed here as Python so they travel with the
package (the GPU box has no reference tree).  `MujocoApp` also accepts the
path of any YAML file with this schema, including the reference's own.

`start_body` note: the shipped default YAMLs leave
`start_body` commented out, which makes `Device.__init__` fail for the arms
(7 chain joints vs 6 start angles).  The built-ins therefore come in two
flavours: `<name>` exactly as shipped and `<name>+start_body`, which adds the
`iros2022.yaml:13,22` setting so the DualUR5 can actually be constructed.
"""
import copy

_K = [1, 2, 3]
_D = [0.5, 1, 1]


def _ctrl(name, kp, kv, ko):
    return {"name": name, "kp": kp, "kv": kv, "ki": 1, "ko": ko, "k": list(_K), "d": list(_D)}


def _devices(arm_abg, base_max_vel, arm_max_vel, base_start, right_start, left_start, start_body):
    arm = lambda idx, name, start: dict(
        id=idx, name=name, max_vel=list(arm_max_vel), EE="ur_EE_" + name,
        ctrlr_dof_xyz=[True, True, True], ctrlr_dof_abg=[arm_abg] * 3,
        start_angles=list(start), num_gripper_joints=6,
        **({"start_body": "dual_ur_stand"} if start_body else {}))
    return [
        dict(id=0, name="base", max_vel=list(base_max_vel), EE="ur_stand_dummy",
             ctrlr_dof_xyz=[False, False, False], ctrlr_dof_abg=[False, False, True],
             start_angles=list(base_start), num_gripper_joints=0),
        arm(1, "ur5right", right_start),
        arm(2, "ur5left", left_start),
    ]


_ROBOTS = [{"id": 0, "name": "DualUR5", "device_ids": [0, 1, 2]}]
_LEFT_DEFAULT = [0.126, -0.942, -1.88, -4.15, -4.78, 0.0]


def _default(arm_abg, osc2_kv, start_body):
    return {
        "devices": _devices(arm_abg, [0, 20], [1, 5], [0.0], [0.0] * 6, _LEFT_DEFAULT, start_body),
        "robots": copy.deepcopy(_ROBOTS),
        "controller_configs": [
            _ctrl("osc0", 2000, 20, 2000), _ctrl("osc1", 200, 50, 200), _ctrl("osc2", 200, osc2_kv, 200),
            {"name": "nullspace", "kv": 10},
        ],
    }


def _iros2022():
    return {
        "devices": _devices(
            True, [0, 2], [2.0, 5], [-1.56],
            [-0.03614821, -0.27430234, -0.47910152, -0.14136462, -0.01368577, -0.54591325],
            [0.05044541, 0.18777629, 0.30305106, -3.10725317, 1.58646237, 2.71767586], True),
        "robots": copy.deepcopy(_ROBOTS),
        "controller_configs": [
            _ctrl("osc0", 200, 20, 75), _ctrl("osc1", 200, 50, 200), _ctrl("osc2", 200, 20, 75),
            {"name": "nullspace", "kv": 10},
        ],
    }


def robot_config(name: str):
    """Fresh (deep-copied) config dict for a built-in name; KeyError if unknown."""
    start_body = name.endswith("+start_body")
    base = name[:-len("+start_body")] if start_body else name
    if base == "default_xyz.yaml":          # gain_test.py:180 (arms position-only, osc2 kv 20)
        return _default(False, 20, start_body)
    if base == "default_xyz_abg.yaml":      # admit_test.py:85, insertion_task.py:424 (osc2 kv 50)
        return _default(True, 50, start_body)
    if base == "iros2022.yaml":
        return _iros2022()
    raise KeyError(name)


# free bodies each demo scene adds after the robot (scene nv = 25 + 6 * count)
SCENE_FREE_OBJECTS = {
    "gain_test_scene.xml": 0,        # gain_test_scene.xml:5-6
    "admit_test_scene.xml": 2,       # admit_test_scene.xml:14-15
    "insertion_task_scene.xml": 4,   # insertion_task_scene.xml:10-13
    "iros2022.xml": 2,               # iros2022.xml:8-9 (quad bracket + quad pegs)
}

# `device_config` block of action_sequence_configs/iros2022_task.yaml:1-4.  The reference ships the
# file but no code reads it (no example loads iros2022_task.yaml; `control_type` appears nowhere in
# its Python), so "joint" vs "task" has no defined behaviour to reproduce: `device_cfgs()` pairs
# devices with controllers in the listed order - which is also the target order - and
# `control_type` is carried along unread, as in the reference.
IROS2022_DEVICE_CONFIG = {
    "devices": ["base", "ur5left", "ur5right"],
    "controllers": ["osc0", "osc2", "osc2"],
    "control_type": ["joint", "task", "task"],
}


def device_cfgs(device_config):
    """[(device name, controller config name)] in the order a `device_config` block lists them."""
    devs, ctrls = device_config["devices"], device_config["controllers"]
    if len(devs) != len(ctrls):
        raise ValueError("device_config: %d devices but %d controllers" % (len(devs), len(ctrls)))
    return list(zip(devs, ctrls))


# ---------------------------------------------------------------- action sequence configs
def _wp(**kw):
    return dict(action="WP", **kw)


def _grip(force, duration):
    return dict(action="GRIP", gripper_force=force, gripper_duration=duration)


def _insertion_task():
    """action_sequence_configs/insertion_task.yaml as Python (same keys and values)."""
    def objects(male_joint, male, female_joint, female):
        return {
            "male_object": dict(joint_name=male_joint, hover_offset=male[0], grip_offset=male[1],
                                initial_pos_xyz=male[2], initial_pos_abg=male[3], grip_yaw=90),
            "female_object": dict(joint_name=female_joint, hover_offset=female[0], hover_offset_lift=female[1],
                                  insert_offset=female[2], initial_pos_xyz=female[3], initial_pos_abg=female[4],
                                  grip_yaw=90),
        }
    return {
        # insertion_task.yaml:1-16
        "grommet_action_objects": objects(
            "free_joint_grommet_11mm", ([0, 0, 0.2], [0, 0, 0.158], [-0.2, 0.5, 0.005], [0, 0, 20]),
            "free_joint_dual_peg", ([0, 0, 0.23], [0, 0, 0.25], [0, 0, 0.18], [-0.4, 0.7, 0.0], [0, 0, 30])),
        # insertion_task.yaml:18-33
        "nist_action_objects": objects(
            "free_joint_male", ([0, 0, 0.32], [0, 0, 0.24], [0.4, 0.6, -0.0515], [0, 0, 30]),
            "free_joint_female", ([0, 0, 0.27], [0, 0, 0.3], [0, 0, 0.22], [0.7, 0.4, -0.00002], [0, 0, -30])),
        # insertion_task.yaml:35-104: the 12 entries
        "insertion_action_sequence": [
            _wp(target_xyz="male_object", target_abg="male_object", offset="hover_offset"),
            _grip(-0.08, 1.0),
            _wp(target_xyz="male_object", target_abg="male_object", offset="grip_offset", gripper_force=-0.08),
            _grip(0.2, 2.0),
            _wp(target_xyz="male_object", target_abg="male_object", gripper_force=0.2, offset="hover_offset"),
            _wp(max_speed_xyz=0.3, target_xyz="female_object", target_abg="male_object", gripper_force=0.2,
                offset="hover_offset"),
            _wp(max_speed_xyz=0.1, target_xyz="female_object", target_abg="female_object", gripper_force=0.2,
                offset="hover_offset"),
            _grip(0.2, 1.0),
            _wp(target_xyz="female_object", target_abg="female_object", max_speed_xyz=1.0, gripper_force=0.1,
                max_error=0.01, offset="insert_offset"),
            _grip(-0.1, 2.0),
            _wp(target_xyz="female_object", target_abg="female_object", gripper_force=-0.1, max_speed_xyz=0.1,
                offset="hover_offset_lift"),
            _wp(target_xyz="start_pos", target_abg="female_object", max_speed_xyz=2.0, gripper_force=0, max_error=0.05),
        ],
    }


def action_config(name: str):
    """Fresh config dict of a built-in action sequence file (`InsertionTask.get_action_config`, insertion_task.py:130-140);
    a path to a YAML file with the same schema is read instead when it exists."""
    import os
    if os.path.isfile(name):
        import yaml
        with open(name, "r") as fh:
            return yaml.safe_load(fh)
    if os.path.basename(name) == "insertion_task.yaml":
        return _insertion_task()
    raise KeyError(name)
