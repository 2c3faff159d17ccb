"""CPU: the host side of the batched insertion demo (`irl_control_b200/insertion.py`: object placement and
`set_waypoint_targets` for B episodes) against the per-episode restatement in oracle/sequence_numpy.py, and the
reference's own 12-entry action list (action_sequence_configs/insertion_task.yaml:35-104) through the state
machine compiled into the fused step (host build) against the restated caller loop."""
import copy
import os

import numpy as np
import pytest

import fused_host
from irl_control_b200 import insertion
from irl_control_b200.configs import action_config
from irl_control_b200.rigid_model import model_for_layout
from irl_control_b200.sequence import ActionSequence, default_ee_quat
from irl_control_b200.synthetic import build_scenario
from oracle import sequence_numpy, t3d
from test_sequence_host import _poses, _trajectory


def test_builtin_action_config_has_the_reference_schema():
    cfg = action_config("insertion_task.yaml")
    seq = cfg["insertion_action_sequence"]
    assert [a["action"] for a in seq] == ["WP", "GRIP", "WP", "GRIP", "WP", "WP", "WP", "GRIP", "WP", "GRIP", "WP", "WP"]
    assert seq[-1]["target_xyz"] == "start_pos" and seq[8]["max_error"] == 0.01 and seq[5]["max_speed_xyz"] == 0.3
    for objs in ("nist_action_objects", "grommet_action_objects"):
        assert set(cfg[objs]) == {"male_object", "female_object"}
        assert cfg[objs]["female_object"]["grip_yaw"] == 90 and len(cfg[objs]["male_object"]["grip_offset"]) == 3
    with pytest.raises(KeyError):
        action_config("no_such_task.yaml")


def test_batched_rotations_match_the_scalar_restatement():
    rng = np.random.default_rng(2)
    e = rng.uniform(-3.1, 3.1, size=(64, 3))
    e[0] = [0.3, np.pi / 2, -0.2]                      # gimbal branch of mat2euler
    q = insertion.euler2quat_b(e)
    m = insertion.euler2mat_b(e)
    for i in range(64):
        assert np.abs(q[i] - t3d.euler2quat(*e[i])).max() < 1e-15
        assert np.abs(m[i] - t3d.euler2mat(*e[i])).max() < 1e-15
        assert np.abs(insertion.quat2mat_b(q)[i] - t3d.quat2mat(q[i])).max() < 1e-15
        assert np.abs(insertion.mat2euler_b(m)[i] - np.array(t3d.mat2euler(m[i]))).max() < 1e-15


@pytest.mark.parametrize("objects_name", ["nist_action_objects", "grommet_action_objects"])
@pytest.mark.parametrize("arm", ["right", "left"])
@pytest.mark.parametrize("randomize", [False, True])
def test_waypoint_poses_match_set_waypoint_targets(objects_name, arm, randomize):
    B = 24
    cfg = action_config("insertion_task.yaml")
    actions, objs = cfg["insertion_action_sequence"], cfg[objects_name]
    rng = np.random.default_rng(5)
    start = rng.uniform(-0.5, 0.5, size=(B, 3)) + [0.3, 0.2, 0.9]
    u = rng.random((B, 6))
    u[0, 4], u[1, 5] = 0.49, 0.51                       # yaw draws of -0.4 and +0.4 degrees: int() gives 0 for both
    placed = insertion.random_object_poses(B, arm, objs, u=u) if randomize else insertion.configured_object_poses(B, objs)
    wp_xyz, wp_quat = insertion.waypoint_poses(actions, objs, placed, start)
    assert wp_xyz.shape == (B, 12, 3) and wp_quat.shape == (B, 12, 4)
    for i in range(B):
        o = copy.deepcopy(objs)
        if randomize:
            draws = iter(u[i])
            qpos = sequence_numpy.initialize_action_objects_random(o, arm, lambda lo, hi: lo + (hi - lo) * next(draws))
        else:
            qpos = sequence_numpy.initialize_action_objects(o)
        for name in qpos:
            assert np.array_equal(placed[name][0][i], qpos[name][0])
            assert np.abs(placed[name][1][i] - qpos[name][1]).max() < 1e-15
        for a, p in enumerate(actions):
            if p["action"] != "WP":
                assert np.array_equal(wp_xyz[i, a], np.zeros(3)) and np.array_equal(wp_quat[i, a], [1, 0, 0, 0])
                continue
            xyz, quat = sequence_numpy.set_waypoint_targets(p, o, qpos, start[i])
            assert np.abs(wp_xyz[i, a] - xyz).max() < 1e-15, (i, a)
            assert np.abs(wp_quat[i, a] - quat).max() < 1e-14, (i, a)
    if randomize:
        sign = 1.0 if arm == "right" else -1.0
        mx = sign * placed["male_object"][0][:, 0]
        assert (mx >= 0.4).all() and (mx <= 0.6).all() and (placed["female_object"][0][:, 1] >= 0.5).all()
        assert np.array_equal(placed["male_object"][1][0], [1.0, 0.0, 0.0, 0.0])      # yaw truncated to 0
        # z keeps the YAML value (only x and y are redrawn, insertion_task.py:359-360)
        assert np.all(placed["male_object"][0][:, 2] == objs["male_object"]["initial_pos_xyz"][2])


def test_waypoint_poses_error_cases():
    cfg = action_config("insertion_task.yaml")
    objs = cfg["nist_action_objects"]
    placed = insertion.configured_object_poses(2, objs)
    start = np.zeros((2, 3))
    with pytest.raises(KeyError):
        insertion.waypoint_poses([{"action": "WP"}], objs, placed, start)
    with pytest.raises(AssertionError):            # list + list = six numbers -> Target.set_xyz asserts (utils.py:36)
        insertion.waypoint_poses([{"action": "WP", "target_xyz": [0.1, 0.2, 0.3]}], objs, placed, start)
    with pytest.raises(ValueError):
        insertion.waypoint_poses([{"action": "WP", "target_xyz": 3.0}], objs, placed, start)
    with pytest.raises(KeyError):
        insertion.waypoint_poses([{"action": "WP", "target_xyz": "no_object"}], objs, placed, start)
    xyz, quat = insertion.waypoint_poses([{"action": "WP", "target_xyz": "start_pos", "target_abg": [0, -90, -90]},
                                          {"action": "WP", "target_xyz": "start_pos"}], objs, placed, start)
    assert np.abs(quat[0, 0] - default_ee_quat()).max() < 1e-15 and np.abs(quat[0, 1] - default_ee_quat()).max() < 1e-15


@pytest.mark.parametrize("active", ["ur5right", "ur5left"])
def test_reference_action_list_through_the_kernel_state_machine(active):
    """All 12 entries of insertion_task.yaml (their kp / speed limits / max_error / gripper forces; GRIP durations
    counted with a 0.25 s control period so that an episode fits in ~100 steps) on the host build of the sequence
    kernel; every step's action index, targets, max_vel[0], error and gripper override equal the restated loop."""
    B, T = 4, 110
    actions = action_config("insertion_task.yaml")["insertion_action_sequence"]
    A = len(actions)
    app, _osc, names, layout = build_scenario("insertion")
    robot = app.get_robot("DualUR5")
    model = model_for_layout(app.sim.model, robot.joint_ids_all, layout)
    passive = [n for n in names if n != active][0]
    seq = ActionSequence(layout, actions, active_arm=active, step_period=0.25)
    assert [p.get("grip_steps") for p in seq.params if p["action"] == "GRIP"] == [4, 8, 4, 8]
    q, dq = _trajectory(B, T, seed=12)
    poses = _poses(layout, q)
    # the waypoint of the i-th WP entry = the pose the active arm will have at a chosen tick
    wp_actions = [a for a, p in enumerate(actions) if p["action"] == "WP"]
    hits = dict(zip(wp_actions, (5, 16, 31, 38, 46, 58, 74, 95)))
    wp_xyz, wp_quat = np.zeros((B, A, 3)), np.zeros((B, A, 4))
    wp_quat[..., 0] = 1.0
    for a, t in hits.items():
        wp_xyz[:, a], wp_quat[:, a] = poses[active][0][t], poses[active][1][t]
    st = seq.new_state(B, wp_xyz, wp_quat)
    mv = np.tile(np.array([list(d.max_vel) for d in layout.devices])[None], (B, 1, 1))
    ia, ip = names.index(active), names.index(passive)
    dev = layout.as_dict()["devices"][ia]
    recs, ctrls = [[] for _ in range(B)], []
    for t in range(T):
        out = fused_host.sequence_step(layout, model, seq, {"q": q[t], "dq": dq[t], "max_vel": mv}, st)
        ctrls.append(out["ctrl"].copy())
        for i in range(B):
            recs[i].append(dict(action=int(st["action"][i]), err=float(st["err"][i]), max_vel0=float(st["max_vel0"][i]),
                                target_xyz=st["target_xyz"][i].copy(), target_quat=st["target_quat"][i].copy()))
    for i in range(B):
        ps = {"active_xyz": poses[active][0][:, i], "active_quat": poses[active][1][:, i], "passive_xyz": poses[passive][0][:, i]}
        ref = sequence_numpy.run_sequence(seq.params, wp_xyz[i], wp_quat[i], ps, dev, default_ee_quat(),
                                          layout.devices[ia].max_vel[0], T)
        for t in range(T):
            r, g = ref[t], recs[i][t]
            assert g["action"] == r["action"], (i, t)
            assert g["max_vel0"] == pytest.approx(r["max_vel0"], rel=1e-12, abs=0), (i, t)
            if r["action"] < A:
                assert (np.isinf(g["err"]) and np.isinf(r["err"])) or g["err"] == pytest.approx(r["err"], rel=1e-9, abs=1e-13)
            assert np.abs(g["target_xyz"][ia] - r["active_xyz"]).max() < 1e-15
            assert np.abs(g["target_quat"][ia] - r["active_quat"]).max() < 1e-15
            assert np.abs(g["target_xyz"][ip] - r["passive_xyz"]).max() < 1e-12
            if r["gripper_force"] != 0.0:
                assert ctrls[t][i, seq.gripper_slot] == r["gripper_force"]
        assert recs[i][-1]["action"] == A, (i, recs[i][-1]["action"])       # all twelve actions completed


# ---------------------------------------------------------------- pinned against the UNMODIFIED reference loop
from oracle import ref_harness  # noqa: E402


def _stream_through(wp_xyz, wp_quat, wp_actions, T, seed):
    """Pose stream of the active arm that passes exactly through the given waypoints at increasing ticks (linear
    approach in between, orientation switching a few ticks before the hit) + a wandering passive arm."""
    rng = np.random.default_rng(seed)
    # leave room for the GRIP budgets between consecutive waypoints
    hits = np.array([8 + 14 * i + int(rng.integers(0, 4)) for i in range(len(wp_actions))])
    assert hits[-1] < T - 2
    xyz = np.zeros((T, 3))
    quat = np.zeros((T, 4))
    prev_t, prev = 0, wp_xyz[wp_actions[0]] + rng.normal(0, 0.3, 3)
    for a, t in zip(wp_actions, hits):
        for s in range(prev_t + (1 if prev_t else 0), t + 1):       # (the previous hit tick keeps its exact pose)
            w = (s - prev_t) / max(1, t - prev_t)
            xyz[s] = (1 - w) * prev + w * wp_xyz[a]
            q = wp_quat[a] + (0.2 * (1 - w)) * rng.normal(size=4)          # noisy until the hit itself
            quat[s] = q / np.linalg.norm(q)
        quat[t] = wp_quat[a]
        prev_t, prev = t, wp_xyz[a]
    xyz[prev_t:], quat[prev_t:] = prev, quat[prev_t]
    passive = np.cumsum(rng.normal(0, 0.01, size=(T, 3)), axis=0) + [-0.4, 0.3, 0.8]
    return {"active_xyz": xyz, "active_quat": quat, "passive_xyz": passive}


@pytest.mark.reference
@pytest.mark.skipif(not ref_harness.reference_available(), reason="needs /root/reference")
@pytest.mark.parametrize("active,objects_name", [("ur5right", "nist_action_objects"), ("ur5left", "grommet_action_objects")])
def test_restated_loop_and_waypoint_poses_match_the_unmodified_reference(active, objects_name):
    """The reference's own `run_sequence`, `go_to_waypoint`, `grip`, `send_forces`, `set_waypoint_targets`
    (insertion_task.py, methods called unmodified on a subclass whose simulator, viewer and timer are stand-ins)
    against (1) `insertion.waypoint_poses` - the product's batched set_waypoint_targets - and (2) the restated loop
    `oracle/sequence_numpy.run_sequence` that the kernel's state machine is checked with."""
    T, dt = 140, 0.25
    cfg = action_config("insertion_task.yaml")
    actions, objs = cfg["insertion_action_sequence"], cfg[objects_name]
    _app, _osc, names, layout = build_scenario("insertion")
    ia = names.index(active)
    placed = insertion.random_object_poses(1, "right" if active == "ur5right" else "left", objs,
                                           rng=np.random.default_rng(4))
    qpos = {objs[k]["joint_name"]: np.concatenate([placed[k][0][0], placed[k][1][0]]) for k in objs}
    wp_actions = [a for a, p in enumerate(actions) if p["action"] == "WP"]
    # start_pos is read from the stream's first tick by both sides; it only matters for the last action, whose
    # waypoint the stream must hit, so fix it first
    start = np.array([0.35, 0.1, 0.85])
    wp_xyz, wp_quat = insertion.waypoint_poses(actions, objs, placed, start[None])
    poses = _stream_through(wp_xyz[0], wp_quat[0], wp_actions, T, seed=1)
    poses["active_xyz"][0] = start
    seq = ActionSequence(layout, actions, active_arm=active, step_period=dt)
    dev = layout.as_dict()["devices"][ia]
    ref = ref_harness.drive_reference_sequence(copy.deepcopy(actions), copy.deepcopy(objs), qpos, poses, active,
                                               dev["ctrlr_dof"], layout.devices[ia].max_vel[0], T, dt)
    mine = sequence_numpy.run_sequence(seq.params, wp_xyz[0], wp_quat[0], poses, dev, default_ee_quat(),
                                       layout.devices[ia].max_vel[0], T)
    # the reference program ends with its sequence; the restated loop keeps holding (action == len) until T
    done = [r for r in mine if r["action"] < len(actions)]
    assert len(ref) == len(done) and len(done) < T, (len(ref), len(done))
    for r, g in zip(ref, mine):
        assert r["tick"] == g["tick"] and r["action"] == g["action"], (r["tick"], r["action"], g["action"])
        assert (np.isinf(r["err"]) and np.isinf(g["err"])) or r["err"] == pytest.approx(g["err"], rel=1e-12, abs=1e-15)
        assert r["max_vel0"] == pytest.approx(g["max_vel0"], rel=1e-12, abs=0)
        assert r["gripper_force"] == g["gripper_force"]
        for k in ("active_xyz", "active_quat", "passive_xyz", "passive_quat"):
            assert np.abs(r[k] - g[k]).max() < 1e-14, (r["tick"], k)
    assert max(r["action"] for r in ref) == len(actions) - 1        # the stream drove the reference through all 12


@pytest.mark.reference
@pytest.mark.skipif(not ref_harness.reference_available(), reason="needs /root/reference")
@pytest.mark.parametrize("objects_name", ["nist_action_objects", "grommet_action_objects"])
def test_object_placement_matches_the_unmodified_reference(objects_name):
    """`insertion.configured_object_poses` / `random_object_poses` against the reference's own
    `initialize_action_objects` / `initialize_action_objects_random` (seeded `np.random`, same draws)."""
    objs = action_config("insertion_task.yaml")[objects_name]
    ref, _ = ref_harness.reference_object_placement(copy.deepcopy(objs))
    mine = insertion.configured_object_poses(1, objs)
    for name, obj in objs.items():
        assert np.abs(mine[name][0][0] - ref[obj["joint_name"]][0]).max() == 0.0
        assert np.abs(mine[name][1][0] - ref[obj["joint_name"]][1]).max() < 1e-15
    lo = np.array([0.4, 0.5, 0.0, 0.5, -20.0, -20.0])
    hi = np.array([0.6, 0.7, 0.3, 0.7, 20.0, 20.0])
    for arm in ("right", "left"):
        for seed in range(12):
            ref, draws = ref_harness.reference_object_placement(copy.deepcopy(objs), arm, seed)
            u = ((np.array(draws) - lo) / (hi - lo))[None]
            mine = insertion.random_object_poses(1, arm, objs, u=u)
            for name, obj in objs.items():
                assert np.abs(mine[name][0][0] - ref[obj["joint_name"]][0]).max() < 1e-15, (arm, seed, name)
                assert np.abs(mine[name][1][0] - ref[obj["joint_name"]][1]).max() < 1e-14, (arm, seed, name)


def test_grip_without_duration_raises_like_the_reference():
    """insertion_task.py:101 spells the default 'gripper_duation', so `params['gripper_duration']` (196) raises
    KeyError for a GRIP entry that does not carry its own duration."""
    _app, _osc, _names, layout = build_scenario("insertion")
    with pytest.raises(KeyError):
        ActionSequence(layout, [{"action": "GRIP", "gripper_force": 0.1}], active_arm="ur5right")
    seq = ActionSequence(layout, [{"action": "GRIP", "gripper_duration": 0.5}], active_arm="ur5right", step_period=0.1)
    assert seq.params[0]["gripper_force"] == -0.08 and seq.params[0]["grip_steps"] == 5


@pytest.mark.parametrize("active", ["ur5right", "ur5left"])
def test_caller_loop_golden_from_the_reference(active):
    """tests/golden/sequence_<arm>.npz: records of the reference's own insertion loop (generated by
    tests/golden/make_golden.py --only-caller-loops from the unmodified methods).  Checked here without the reference:
    the product's object placement and waypoint poses, and the restated loop the kernel is compared with."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sequence_%s.npz" % active))
    cfg = action_config("insertion_task.yaml")
    actions, objs = cfg["insertion_action_sequence"], cfg[str(g["objects_name"])]
    lo = np.array([0.4, 0.5, 0.0, 0.5, -20.0, -20.0])
    hi = np.array([0.6, 0.7, 0.3, 0.7, 20.0, 20.0])
    placed = insertion.random_object_poses(1, str(g["arm"]), objs, u=((g["draws"] - lo) / (hi - lo))[None])
    for name, key in (("male_object", "male_qpos"), ("female_object", "female_qpos")):
        assert np.abs(placed[name][0][0] - g[key][:3]).max() < 1e-15 and np.abs(placed[name][1][0] - g[key][3:]).max() < 1e-14
    wp_xyz, wp_quat = insertion.waypoint_poses(actions, objs, placed, g["start_pos"][None])
    _app, _osc, names, layout = build_scenario("insertion")
    ia = names.index(active)
    seq = ActionSequence(layout, actions, active_arm=active, step_period=float(g["step_period"]))
    poses = {k: g["pose_" + k] for k in ("active_xyz", "active_quat", "passive_xyz")}
    T = int(g["n_ticks"])
    mine = sequence_numpy.run_sequence(seq.params, wp_xyz[0], wp_quat[0], poses, layout.as_dict()["devices"][ia],
                                       default_ee_quat(), float(g["max_vel0_initial"]), T)
    n = len(g["tick"])
    assert n < T and [r["action"] for r in mine[n:]] == [len(actions)] * (len(mine) - n)
    for t in range(n):
        r = mine[t]
        assert r["tick"] == int(g["tick"][t]) and r["action"] == int(g["action"][t]), t
        assert (np.isinf(r["err"]) and np.isinf(g["err"][t])) or r["err"] == pytest.approx(float(g["err"][t]), rel=1e-12, abs=1e-15)
        assert r["max_vel0"] == pytest.approx(float(g["max_vel0"][t]), rel=1e-12, abs=0)
        assert r["gripper_force"] == float(g["gripper_force"][t])
        for k in ("active_xyz", "active_quat", "passive_xyz", "passive_quat"):
            assert np.abs(r[k] - g[k][t]).max() < 1e-14, (t, k)
    assert int(g["action"].max()) == len(actions) - 1
