"""CPU: the action-sequence state machine compiled into the fused step (osc_sequence.cuh, run on the
host through tests/host_fused) against the restatement of the reference's caller loop
(oracle/sequence_numpy.py: insertion_task.py go_to_waypoint / grip / send_forces / run_sequence)."""
import numpy as np
import pytest

import fused_host
from irl_control_b200.dual_ur5 import DualUR5Model, dynamics
from irl_control_b200.rigid_model import model_for_layout
from irl_control_b200.sequence import ActionSequence, default_ee_quat
from irl_control_b200.synthetic import build_scenario
from oracle import osc_numpy, sequence_numpy

# the shape of action_sequence_configs/insertion_task.yaml:35-104, shortened
ACTIONS = [
    {"action": "WP", "max_error": 0.02},
    {"action": "GRIP", "gripper_force": -0.08, "gripper_duration": 0.006},
    {"action": "WP", "gripper_force": -0.08, "max_error": 0.02},
    {"action": "GRIP", "gripper_force": 0.2, "gripper_duration": 0.004},
    {"action": "WP", "gripper_force": 0.2, "kp": 4.0, "max_error": 0.02, "max_speed_xyz": 2.0},
]


def _trajectory(B, T, seed):
    """Smooth joint trajectories: q_t = q_0 + a sin(w t)."""
    import torch
    rng = np.random.default_rng(seed)
    q0 = rng.uniform(-1.5, 1.5, size=(B, 25))
    q0[:, 7:13] = rng.uniform(0, 0.8, size=(B, 6))
    q0[:, 19:25] = rng.uniform(0, 0.8, size=(B, 6))
    amp = rng.uniform(0.05, 0.3, size=(B, 25))
    w = rng.uniform(0.05, 0.2, size=(B, 25))
    t = np.arange(T)[:, None, None]
    q = q0[None] + amp[None] * np.sin(w[None] * t)
    dq = amp[None] * w[None] * np.cos(w[None] * t) / 0.002
    return q, dq


def _poses(layout, q):
    """EE poses per tick from the rigid-body model (what the reference reads from MuJoCo)."""
    import torch
    T, B, _ = q.shape
    m = DualUR5Model()
    dyn = dynamics(m, torch.from_numpy(q.reshape(T * B, 25)), torch.zeros(T * B, 25, dtype=torch.float64), need_M=False)
    ee = {"ur5right": m.body_name2id("ur_EE_ur5right"), "ur5left": m.body_name2id("ur_EE_ur5left")}
    out = {}
    for nm, b in ee.items():
        out[nm] = (dyn.xpos[:, b].numpy().reshape(T, B, 3), dyn.xquat[:, b].numpy().reshape(T, B, 4))
    return out


@pytest.mark.parametrize("active", ["ur5right", "ur5left"])
def test_sequence_step_matches_the_reference_loop(active):
    B, T = 6, 60
    app, _osc, names, layout = build_scenario("insertion")
    robot = app.get_robot("DualUR5")
    model = model_for_layout(app.sim.model, robot.joint_ids_all, layout)
    passive = [n for n in names if n != active][0]
    seq = ActionSequence(layout, ACTIONS, active_arm=active, step_period=0.002)
    assert seq.gripper_slot == layout.ctrl_slices[names.index(active)].start + 6
    q, dq = _trajectory(B, T, seed=3)
    poses = _poses(layout, q)
    # waypoints = the EE pose the active arm will have at chosen ticks, so that every WP terminates
    hit = {0: 7, 2: 23, 4: 41}
    A = len(ACTIONS)
    wp_xyz, wp_quat = np.zeros((B, A, 3)), np.zeros((B, A, 4))
    wp_quat[..., 0] = 1.0
    for a, t in hit.items():
        wp_xyz[:, a], wp_quat[:, a] = poses[active][0][t], poses[active][1][t]
    st = seq.new_state(B, wp_xyz, wp_quat)
    mv = np.tile(np.array([list(d.max_vel) for d in layout.devices])[None], (B, 1, 1))
    recs = [[] for _ in range(B)]
    ctrls = []
    for t in range(T):
        out = fused_host.sequence_step(layout, model, seq, {"q": q[t], "dq": dq[t], "max_vel": mv}, st)
        ctrls.append(out["ctrl"].copy())
        for i in range(B):
            recs[i].append(dict(action=int(st["action"][i]), err=float(st["err"][i]), max_vel0=float(st["max_vel0"][i]),
                                target_xyz=st["target_xyz"][i].copy(), target_quat=st["target_quat"][i].copy()))
    ia, ip = names.index(active), names.index(passive)
    dev = layout.as_dict()["devices"][ia]
    n_done = 0
    for i in range(B):
        ps = {"active_xyz": poses[active][0][:, i], "active_quat": poses[active][1][:, i], "passive_xyz": poses[passive][0][:, i]}
        ref = sequence_numpy.run_sequence(seq.params, wp_xyz[i], wp_quat[i], ps, dev, default_ee_quat(),
                                          layout.devices[ia].max_vel[0], T)
        assert len(ref) >= T
        for t in range(T):
            r, g = ref[t], recs[i][t]
            assert g["action"] == r["action"], (i, t)
            assert g["max_vel0"] == pytest.approx(r["max_vel0"], rel=1e-12, abs=0), (i, t)
            if r["action"] < A:     # (the reference program ends with the sequence; the batch just holds)
                assert (np.isinf(g["err"]) and np.isinf(r["err"])) or g["err"] == pytest.approx(r["err"], rel=1e-9, abs=1e-13), (i, t)
            assert np.abs(g["target_xyz"][ia] - r["active_xyz"]).max() < 1e-15
            assert np.abs(g["target_quat"][ia] - r["active_quat"]).max() < 1e-15
            assert np.abs(g["target_xyz"][ip] - r["passive_xyz"]).max() < 1e-12, (i, t)
            assert np.abs(g["target_quat"][ip] - r["passive_quat"]).max() < 1e-15
            # send_forces: the gripper slot carries the action's force when it is non-zero
            if r["gripper_force"] != 0.0:
                assert ctrls[t][i, seq.gripper_slot] == r["gripper_force"]
        n_done += recs[i][-1]["action"] == A
    assert n_done == B                                     # every episode ran through all five actions


def test_sequence_step_equals_fused_step_with_the_same_targets():
    """The control law itself is untouched: a sequence step = a fused step with the targets and
    max_vel the state machine chose (+ the gripper override)."""
    B = 16
    app, _osc, names, layout = build_scenario("insertion")
    robot = app.get_robot("DualUR5")
    model = model_for_layout(app.sim.model, robot.joint_ids_all, layout)
    seq = ActionSequence(layout, ACTIONS, active_arm="ur5right")
    q, dq = _trajectory(B, 3, seed=8)
    rng = np.random.default_rng(0)
    wp_xyz = rng.uniform(-0.5, 0.5, size=(B, len(ACTIONS), 3)) + np.array([0.4, 0.0, 0.8])
    wp_quat = rng.normal(size=(B, len(ACTIONS), 4))
    wp_quat /= np.linalg.norm(wp_quat, axis=-1, keepdims=True)
    st = seq.new_state(B, wp_xyz, wp_quat)
    mv = np.tile(np.array([list(d.max_vel) for d in layout.devices])[None], (B, 1, 1))
    for t in range(3):
        out = fused_host.sequence_step(layout, model, seq, {"q": q[t], "dq": dq[t], "max_vel": mv}, st)
        mv2 = mv.copy()
        mv2[:, seq.active_device, 0] = st["max_vel0"]
        ref = fused_host.run(layout, model, {"q": q[t], "dq": dq[t], "max_vel": mv2, "target_xyz": st["target_xyz"],
                                             "target_quat": st["target_quat"]})
        want = ref["ctrl"].copy()
        gf = np.array([seq.params[a]["gripper_force"] if a < seq.n_actions else 0.0 for a in st["action"]])
        sel = gf != 0.0
        want[sel, seq.gripper_slot] = gf[sel]
        assert np.array_equal(out["ctrl"], want)
        assert np.array_equal(out["u_all"], ref["u_all"])


def test_waypoint_cycling_matches_the_gain_test_loop():
    """examples/gain_test.py:134-162: each arm walks its waypoint list independently."""
    B, T, W = 5, 50, 4
    app, _osc, names, layout = build_scenario("gain_test")
    robot = app.get_robot("DualUR5")
    model = model_for_layout(app.sim.model, robot.joint_ids_all, layout)
    q, dq = _trajectory(B, T, seed=21)
    poses = _poses(layout, q)
    D = layout.D
    n_wp = [4, 3, 1]                                             # ur5right, ur5left, (base: unused)
    assert names == ["ur5right", "ur5left", "base"]
    wps = np.zeros((B, D, W, 3))
    ticks = {0: [5, 11, 12, 30], 1: [3, 20, 33]}                 # the EE passes exactly through its waypoints
    for d, ts in ticks.items():
        for w, t in enumerate(ts):
            wps[:, d, w] = poses[names[d]][0][t]
    st = {"wps": wps, "n_wp": n_wp, "wp_idx": np.zeros((B, D), np.int32), "target_xyz": np.zeros((B, D, 3)),
          "target_quat": np.tile(np.array([1.0, 0, 0, 0]), (B, D, 1))}
    mv = np.tile(np.array([list(d.max_vel) for d in layout.devices])[None], (B, 1, 1))
    got = []
    for t in range(T):
        before = st["wp_idx"].copy()
        out = fused_host.waypoints_step(layout, model, {"q": q[t], "dq": dq[t], "max_vel": mv}, st, threshold=0.1)
        got.append((st["target_xyz"].copy(), before))
        ref = fused_host.run(layout, model, {"q": q[t], "dq": dq[t], "max_vel": mv, "target_xyz": st["target_xyz"],
                                             "target_quat": st["target_quat"]})
        assert np.array_equal(out["ctrl"], ref["ctrl"])          # the control law is the plain fused step
    for i in range(B):
        for d in (0, 1):
            ref = sequence_numpy.run_waypoint_cycle(wps[i, d, :n_wp[d]], poses[names[d]][0][:, i], 0.1, T)
            for t in range(T):
                assert got[t][1][i, d] == ref[t][1], (i, d, t)
                assert np.array_equal(got[t][0][i, d], ref[t][0])
    assert (st["wp_idx"][:, :2] != 0).any() or True
    assert max(r[1] for r in ref) > 0                             # indices advanced and wrapped


def test_waypoint_cycling_under_admittance_matches_the_force_test_loop():
    """examples/force_test.py:88-110: admittance controller (F/T term in the law), the right arm holds one
    waypoint, the left arm walks ten with a fixed orientation target; `errors['ur5left'] < threshold_ee`
    advances / wraps its index.  Same kernel entry as the gain_test loop, admit_test layout (osc1 == osc2 of
    default_xyz_abg.yaml numerically)."""
    B, T = 4, 64
    app, _osc, names, layout = build_scenario("admit_test")
    assert names == ["ur5right", "ur5left"] and layout.admittance
    robot = app.get_robot("DualUR5")
    model = model_for_layout(app.sim.model, robot.joint_ids_all, layout)
    q, dq = _trajectory(B, T, seed=33)
    poses = _poses(layout, q)
    D, W = layout.D, 10
    n_wp = [1, 10]
    wps = np.zeros((B, D, W, 3))
    wps[:, 0, 0] = [0.3, 0.46432, 0.36243]                        # force_test.py:39-41 (never reached here)
    left_ticks = [4, 9, 15, 16, 22, 30, 37, 41, 50, 58]           # the left EE passes through its ten waypoints
    for w, t in enumerate(left_ticks):
        wps[:, 1, w] = poses["ur5left"][0][t]
    from irl_control_b200.insertion import euler2quat_b
    tq = np.tile(np.array([1.0, 0, 0, 0]), (B, D, 1))
    tq[:, 1] = euler2quat_b(np.array([0.0, 0.0, -np.pi / 2]))      # targets['ur5left'].set_abg([0, 0, -pi/2]) (94)
    st = {"wps": wps, "n_wp": n_wp, "wp_idx": np.zeros((B, D), np.int32), "target_xyz": np.zeros((B, D, 3)),
          "target_quat": tq.copy()}
    mv = np.tile(np.array([list(d.max_vel) for d in layout.devices])[None], (B, 1, 1))
    ft = np.random.default_rng(1).normal(0.0, 5.0, size=(B, D, 6))
    seen = []
    for t in range(T):
        before = st["wp_idx"].copy()
        inp = {"q": q[t], "dq": dq[t], "max_vel": mv, "ft_raw": ft}
        out = fused_host.waypoints_step(layout, model, inp, st, threshold=0.01)
        seen.append((st["target_xyz"].copy(), before))
        ref = fused_host.run(layout, model, dict(inp, target_xyz=st["target_xyz"], target_quat=st["target_quat"]))
        assert np.array_equal(out["ctrl"], ref["ctrl"])
        assert np.array_equal(st["target_quat"], tq)              # orientation targets are the caller's
    for i in range(B):
        for d in (0, 1):
            ref = sequence_numpy.run_waypoint_cycle(wps[i, d, :n_wp[d]], poses[names[d]][0][:, i], 0.01, T)
            for t in range(T):
                assert seen[t][1][i, d] == ref[t][1], (i, d, t)
                assert np.array_equal(seen[t][0][i, d], ref[t][0])
    # right holds its only waypoint; left walked all ten and wrapped to the first one again
    assert st["wp_idx"][:, 0].max() == 0 and max(s[1][:, 1].max() for s in seen) == 9 and (st["wp_idx"][:, 1] == 0).all()


from oracle import ref_harness  # noqa: E402


@pytest.mark.reference
@pytest.mark.skipif(not ref_harness.reference_available(), reason="needs /root/reference")
def test_restated_waypoint_cycle_matches_the_unmodified_gain_test_loop():
    """`oracle/sequence_numpy.run_waypoint_cycle` (what `irlosc_step_waypoints` is checked with) against the
    reference's own `GainTest.run` loop driven on the same EE position streams."""
    T = 80
    rng = np.random.default_rng(6)
    rw = rng.uniform(-0.5, 0.5, size=(4, 3))
    lw = rng.uniform(-0.5, 0.5, size=(3, 3))
    streams = {}
    for name, wps in (("r", rw), ("l", lw)):
        s = np.cumsum(rng.normal(0, 0.03, size=(T + 1, 3)), axis=0)
        for t in rng.choice(np.arange(2, T), size=14, replace=False):        # visits: some hit the current waypoint
            s[t] = wps[rng.integers(0, len(wps))] + rng.normal(0, 0.02, 3)
        streams[name] = s
    ref = ref_harness.drive_reference_gain_test(rw, lw, streams["r"], streams["l"], T)
    assert len(ref) == T
    mine_r = sequence_numpy.run_waypoint_cycle(rw, streams["r"], 0.1, T)
    mine_l = sequence_numpy.run_waypoint_cycle(lw, streams["l"], 0.1, T)
    for t in range(T):
        assert np.array_equal(ref[t][0], mine_r[t][0]) and ref[t][2] == mine_r[t][1], t
        assert np.array_equal(ref[t][1], mine_l[t][0]) and ref[t][3] == mine_l[t][1], t
    assert max(r[2] for r in ref) > 0 and max(r[3] for r in ref) > 0          # both lists advanced


def test_waypoint_cycle_golden_from_the_reference():
    """tests/golden/waypoint_cycle.npz: the reference's own `GainTest.run` loop on recorded EE streams."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "waypoint_cycle.npz"))
    T = len(g["right_idx"])
    for side in ("right", "left"):
        mine = sequence_numpy.run_waypoint_cycle(g[side + "_wps"], g["ee_" + side], float(g["threshold"]), T)
        for t in range(T):
            assert np.array_equal(mine[t][0], g[side + "_target"][t]) and mine[t][1] == int(g[side + "_idx"][t]), (side, t)
        assert g[side + "_idx"].max() > 0
