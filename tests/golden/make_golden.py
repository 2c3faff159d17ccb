#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference `OSC.generate`.

Run in the build container (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

For every scenario of SURVEY.md section 8 (gain_test, admit_test, insertion,
worst_case k = 13) plus the quirk cases (non-zero target velocity N3/N4,
near-singular start pose N2, per-instance max_vel schedule) it draws seeded
states with `irl_control_b200.synthetic.synth_batch`, feeds each instance to
the reference's own Device / Robot / OSC classes through oracle/ref_harness.py
(stub mujoco_py / transforms3d, fake sim) and stores inputs + outputs:

    inputs : M, J6, dq, bias, ee_xyz, ee_quat, ft_xmat, ft_raw, target_xyz,
             target_quat, target_vel, max_vel            (per-device arrays in target order)
    outputs: ctrl (packed forces, target order), u_all (n), pinv (branch flag)
    meta   : layout_json (the flattened controller description), scenario, seed
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from irl_control_b200.dual_ur5 import DualUR5Model  # noqa: E402
from irl_control_b200.synthetic import SCENARIOS, build_scenario, synth_batch  # noqa: E402
from irl_control_b200.configs import SCENE_FREE_OBJECTS  # noqa: E402
from oracle import ref_harness  # noqa: E402


def reference_runner(scenario: str, use_g: bool = True, nullspace: bool = True):
    sc = SCENARIOS[scenario]
    yaml_name = sc["config"].replace("+start_body", "")
    cfg = ref_harness.load_reference_yaml(yaml_name, inject_start_body=True)
    for dev in cfg["devices"]:                          # a scenario may stand for a user-edited YAML
        dev.update(sc.get("config_patch", {}).get(dev["name"], {}))
    model = DualUR5Model(n_free_objects=SCENE_FREE_OBJECTS[sc["scene"]])
    return ref_harness.ReferenceRunner(model, cfg, sc["device_cfgs"], sc["targets"], "nullspace" if nullspace else None,
                                       use_g=use_g, admittance=sc["admittance"])


def run_case(name, scenario, B, seed, mutate=None, insertion_schedule=False, no_max_vel=(), use_g=True,
             nullspace=True):
    """no_max_vel: device names whose `max_vel` is cleared to None before generate (osc.py:163-168);
    use_g / nullspace: the OSC constructor's `use_g` flag and `nullspace_config=None` (osc.py:19,190,195)."""
    import dataclasses
    _, _, targets, layout = build_scenario(scenario)
    if not use_g or not nullspace:
        layout = dataclasses.replace(layout, use_g=use_g, nullspace_kv=layout.nullspace_kv if nullspace else None)
    if no_max_vel:
        layout = dataclasses.replace(layout, devices=tuple(
            dataclasses.replace(d, has_max_vel=False) if d.name in no_max_vel else d for d in layout.devices))
    st = synth_batch(layout, B, seed=seed, insertion_schedule=insertion_schedule)
    st = {k: v.numpy() for k, v in st.items()}
    st["target_vel"] = np.zeros((B, layout.D, 6))
    if mutate is not None:
        mutate(st, layout)
    runner = reference_runner(scenario, use_g=use_g, nullspace=nullspace)
    # the host layer's index maps must be the reference's own
    for dl in layout.devices:
        ref_dev = runner.robot.get_device(dl.name)
        assert list(ref_dev.joint_ids_all) == list(dl.joint_ids_all), dl.name
        assert list(ref_dev.actuator_trnids) == list(dl.actuator_trnids), dl.name
        assert list(ref_dev.ctrl_idxs) == list(dl.ctrl_idxs), dl.name
    ctrl, u_all, pinv, raised = [], [], [], []
    for i in range(B):
        inst = {k: v[i] for k, v in st.items()}
        tv = st["target_vel"][i] if np.any(st["target_vel"][i] != 0) else None
        mv = np.array(st["max_vel"][i], dtype=np.float64)
        for d, dl in enumerate(layout.devices):
            if dl.name in no_max_vel:
                mv[d] = np.nan                      # ReferenceRunner.run: device.max_vel = None
        try:
            r = runner.run(inst, st["target_xyz"][i], st["target_quat"][i], tgt_vel=tv, max_vel=mv)
            ctrl.append(np.concatenate(r["forces"]))
            u_all.append(r["u_all"])
            pinv.append(r["pinv"])
            raised.append(False)
            st["target_vel"][i] = r["target_vel_seen"]
        except IndexError:
            ctrl.append(np.full(layout.n_ctrl, np.nan))
            u_all.append(np.full(layout.n, np.nan))
            pinv.append(False)
            raised.append(True)
    out = dict(st)
    out.pop("J")
    out.update(ctrl=np.stack(ctrl), u_all=np.stack(u_all), pinv=np.array(pinv), index_error=np.array(raised),
               layout_json=np.array(json.dumps(layout.as_dict())), scenario=np.array(scenario), seed=np.array(seed))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-28s B=%3d k=%2d pinv=%2d index_error=%2d  max|ctrl|=%.3e -> %s" % (
        name, B, layout.k, int(np.sum(pinv)), int(np.sum(raised)), np.nanmax(np.abs(out["ctrl"])),
        os.path.relpath(path, ROOT)))


def vel_all_nonzero(st, layout):
    rng = np.random.default_rng(5)
    B = st["dq"].shape[0]
    tv = rng.normal(0.0, 0.2, size=(B, layout.D, 6))
    tv[np.abs(tv) < 1e-3] = 0.05
    # odd instances: one zero component on device 0 -> that device falls back to the zero branch (N4)
    tv[1::2, 0, 4] = 0.0
    st["target_vel"] = tv


def singular_pose(st, layout):
    """Right arm at its all-zero start angles (default_xyz_abg.yaml:17): kinematic singularity (N2)."""
    from irl_control_b200.dual_ur5 import dynamics
    model = DualUR5Model()
    B = st["dq"].shape[0]
    rng = np.random.default_rng(11)
    q = rng.uniform(-np.pi, np.pi, size=(B, 25))
    q[:, 1:7] = rng.normal(0.0, 1e-3, size=(B, 6)) * np.linspace(0, 1, B)[:, None]
    q[:, 7:13] = 0.0
    q[:, 19:25] = 0.0
    dyn = dynamics(model, torch.from_numpy(q), torch.from_numpy(st["dq"]))
    ee = {"base": "ur_stand_dummy", "ur5right": "ur_EE_ur5right", "ur5left": "ur_EE_ur5left"}
    st["M"] = dyn.M.numpy()
    st["bias"] = dyn.bias.numpy()
    for d, dl in enumerate(layout.devices):
        b = model.body_name2id(ee[dl.name])
        jp, jr = dyn.jac_body(b)
        st["J6"][:, d] = torch.cat([jp, jr], 1).numpy()
        delta = st["target_xyz"][:, d] - st["ee_xyz"][:, d]
        st["ee_xyz"][:, d] = dyn.xpos[:, b].numpy()
        st["ee_quat"][:, d] = dyn.xquat[:, b].numpy()
        st["target_xyz"][:, d] = st["ee_xyz"][:, d] + delta


def no_max_vel_cases():
    """osc.py:163-168, the branch no shipped YAML takes: `device.max_vel is None` -> gains x stiffness, no limiter."""
    run_case("gain_test_nomaxvel_s10", "gain_test", 12, 10, no_max_vel=("ur5left", "base"))
    run_case("admit_nomaxvel_s11", "admit_test", 12, 11, no_max_vel=("ur5right",))
    # constructor options no example uses: OSC(..., use_g=False) and OSC(..., nullspace_config=None)
    run_case("gain_test_no_g_s12", "gain_test", 12, 12, use_g=False)
    run_case("admit_no_nullspace_s13", "admit_test", 12, 13, nullspace=False)
    run_case("worst_case_bare_s14", "worst_case", 12, 14, use_g=False, nullspace=False)
    # DoF masks no shipped YAML has (5 + 4 + 1 rows): the generic kernel's case
    run_case("mixed_dof_s15", "mixed_dof", 12, 15)
    run_case("mixed_dof_vel_s16", "mixed_dof", 8, 16, mutate=vel_all_nonzero)
    # admittance with the base among the targets
    run_case("worst_case_admit_s17", "worst_case_admit", 12, 17, mutate=base_without_sensor)


def base_without_sensor(st, layout):
    """The base has no F/T sensor: what the state pull hands over for it is a zero wrench in an identity frame."""
    d = [dl.name for dl in layout.devices].index("base")
    st["ft_raw"][:, d] = 0.0
    st["ft_xmat"][:, d] = np.eye(3).reshape(-1)


def iros2022_cases():
    """iros2022.yaml gains / max_vel with the device order of iros2022_task.yaml:1-4."""
    run_case("iros2022_s8", "iros2022", 24, 8)
    run_case("iros2022_vel_s9", "iros2022", 12, 9, mutate=vel_all_nonzero)


def caller_loop_goldens():
    """The reference's caller loops (examples/insertion_task.py, examples/gain_test.py), their own unmodified methods
    driven on pose streams through oracle/ref_harness: one record per `controller.generate` call."""
    import copy
    sys.path.insert(0, os.path.dirname(HERE))
    from test_insertion_host import _stream_through
    from irl_control_b200 import insertion
    from irl_control_b200.configs import action_config
    T, dt = 140, 0.25
    cfg = action_config("insertion_task.yaml")
    actions = cfg["insertion_action_sequence"]
    _, _, names, layout = build_scenario("insertion")
    for active, objects_name, seed in (("ur5right", "nist_action_objects", 21), ("ur5left", "grommet_action_objects", 22)):
        objs = cfg[objects_name]
        arm = "right" if active == "ur5right" else "left"
        ref_placed, draws = ref_harness.reference_object_placement(copy.deepcopy(objs), arm, seed)
        qpos = {jn: np.concatenate([p, q]) for jn, (p, q) in ref_placed.items()}
        placed = {k: (qpos[objs[k]["joint_name"]][None, :3], qpos[objs[k]["joint_name"]][None, 3:]) for k in objs}
        start = np.array([0.35 if arm == "right" else -0.35, 0.1, 0.85])
        wp_xyz, wp_quat = insertion.waypoint_poses(actions, objs, placed, start[None])     # only shapes the stream
        wp_actions = [a for a, p in enumerate(actions) if p["action"] == "WP"]
        poses = _stream_through(wp_xyz[0], wp_quat[0], wp_actions, T, seed=seed)
        poses["active_xyz"][0] = start
        ia = names.index(active)
        dof = layout.as_dict()["devices"][ia]["ctrlr_dof"]
        mv0 = layout.devices[ia].max_vel[0]
        rec = ref_harness.drive_reference_sequence(copy.deepcopy(actions), copy.deepcopy(objs), qpos, poses, active,
                                                   dof, mv0, T, dt)
        assert max(r["action"] for r in rec) == len(actions) - 1
        out = {k: np.stack([np.asarray(r[k], dtype=np.float64) for r in rec]) for k in
               ("tick", "action", "err", "max_vel0", "gripper_force", "active_xyz", "active_quat", "passive_xyz", "passive_quat")}
        out.update({"pose_" + k: v for k, v in poses.items()})
        out.update(draws=np.array(draws), male_qpos=qpos[objs["male_object"]["joint_name"]],
                   female_qpos=qpos[objs["female_object"]["joint_name"]], start_pos=start, step_period=np.array(dt),
                   n_ticks=np.array(T), max_vel0_initial=np.array(mv0), active=np.array(active),
                   objects_name=np.array(objects_name), arm=np.array(arm))
        path = os.path.join(HERE, "sequence_%s.npz" % active)
        np.savez_compressed(path, **out)
        print("%-28s %3d generate() calls through %d actions -> %s" % ("sequence_" + active, len(rec), len(actions),
                                                                       os.path.relpath(path, ROOT)))
    # gain_test waypoint cycling (gain_test.py:134-162)
    T = 80
    rng = np.random.default_rng(6)
    rw, lw = rng.uniform(-0.5, 0.5, size=(4, 3)), rng.uniform(-0.5, 0.5, size=(3, 3))
    streams = {}
    for name, wps in (("r", rw), ("l", lw)):
        s = np.cumsum(rng.normal(0, 0.03, size=(T + 1, 3)), axis=0)
        for t in rng.choice(np.arange(2, T), size=14, replace=False):
            s[t] = wps[rng.integers(0, len(wps))] + rng.normal(0, 0.02, 3)
        streams[name] = s
    rec = ref_harness.drive_reference_gain_test(rw, lw, streams["r"], streams["l"], T)
    path = os.path.join(HERE, "waypoint_cycle.npz")
    np.savez_compressed(path, right_wps=rw, left_wps=lw, ee_right=streams["r"], ee_left=streams["l"],
                        right_target=np.stack([r[0] for r in rec]), left_target=np.stack([r[1] for r in rec]),
                        right_idx=np.array([r[2] for r in rec]), left_idx=np.array([r[3] for r in rec]),
                        threshold=np.array(0.1))
    print("%-28s %3d generate() calls -> %s" % ("waypoint_cycle", len(rec), os.path.relpath(path, ROOT)))


if __name__ == "__main__":
    assert ref_harness.reference_available(), "needs /root/reference"
    if "--only-caller-loops" in sys.argv:      # added after the OSC.generate files were committed
        caller_loop_goldens()
        sys.exit(0)
    if "--only-no-max-vel" in sys.argv:
        no_max_vel_cases()
        sys.exit(0)
    if "--only-iros2022" in sys.argv:          # added after the first eight files were committed
        iros2022_cases()
        sys.exit(0)
    run_case("gain_test_s0", "gain_test", 24, 0)
    run_case("admit_test_s1", "admit_test", 24, 1)
    run_case("insertion_s2", "insertion", 24, 2, insertion_schedule=True)
    run_case("worst_case_s3", "worst_case", 24, 3)
    run_case("gain_test_vel_s4", "gain_test", 12, 4, mutate=vel_all_nonzero)
    run_case("worst_case_vel_s5", "worst_case", 12, 5, mutate=vel_all_nonzero)
    run_case("insertion_vel_s6", "insertion", 8, 6, mutate=vel_all_nonzero)
    run_case("admit_singular_s7", "admit_test", 16, 7, mutate=singular_pose)
    iros2022_cases()
    no_max_vel_cases()
    caller_loop_goldens()
