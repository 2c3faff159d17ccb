"""GPU: `irlosc_step_sequence` (action-sequence state machine inside the fused step) against the
restatement of the reference's caller loop and against plain fused steps with the same targets."""
import numpy as np
import pytest

from test_sequence_host import ACTIONS, _poses, _trajectory

pytestmark = pytest.mark.gpu


def test_sequence_on_gpu_matches_reference_loop_and_fused_step():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from irl_control_b200.engine import BatchedOSC
    from irl_control_b200.sequence import ActionSequence, default_ee_quat
    from irl_control_b200.synthetic import scenario_model
    from oracle import sequence_numpy
    B, T = 1024, 60
    dev = "cuda:0"
    layout, model = scenario_model("insertion")
    names = [d.name for d in layout.devices]
    for active in ("ur5right", "ur5left"):
        passive = [n for n in names if n != active][0]
        ia, ip = names.index(active), names.index(passive)
        eng = BatchedOSC(layout, device=0)
        eng.set_model(model)
        seq = ActionSequence(layout, ACTIONS, active_arm=active, step_period=0.002)
        q, dq = _trajectory(B, T, seed=17)
        n_chk = 8
        poses = _poses(layout, q[:, :n_chk])
        A = len(ACTIONS)
        rng = np.random.default_rng(5)
        wp_xyz = rng.uniform(-0.4, 0.4, size=(B, A, 3)) + np.array([0.3, 0.0, 0.9])     # mostly unreachable in 60 ticks
        wp_quat = rng.normal(size=(B, A, 4))
        wp_quat /= np.linalg.norm(wp_quat, axis=-1, keepdims=True)
        for a, t in {0: 7, 2: 23, 4: 41}.items():                                         # the checked ones terminate
            wp_xyz[:n_chk, a], wp_quat[:n_chk, a] = poses[active][0][t], poses[active][1][t]
        st = seq.new_state(B, wp_xyz, wp_quat, device=dev)
        mv = torch.tensor([list(d.max_vel) for d in layout.devices], dtype=torch.float64, device=dev)[None].expand(B, -1, -1).contiguous()
        qd, dqd = torch.from_numpy(q).to(dev), torch.from_numpy(dq).to(dev)
        recs = []
        for t in range(T):
            out = eng.step_sequence({"q": qd[t].contiguous(), "dq": dqd[t].contiguous(), "max_vel": mv}, seq, st, want_u_all=True)
            # same step without the state machine: identical targets / max_vel -> identical torques
            mv2 = mv.clone()
            mv2[:, seq.active_device, 0] = st["max_vel0"]
            ref = eng.step_fused({"q": qd[t].contiguous(), "dq": dqd[t].contiguous(), "max_vel": mv2,
                                  "target_xyz": st["target_xyz"], "target_quat": st["target_quat"]}, want_u_all=True)
            # (another kernel - the plain step of a small batch runs a lane per arm - and the active arm is processed
            #  first: equal up to rounding amplified by the conditioning of the task-space matrix)
            scale = ref["u_all"].abs().amax(dim=1, keepdim=True)
            worst = ((out["u_all"] - ref["u_all"]).abs() / scale).max().item()
            assert worst < 1e-6, worst
            gf = torch.tensor([p["gripper_force"] for p in seq.params] + [0.0], dtype=torch.float64, device=dev)[st["action"].long()]
            want = ref["ctrl"].clone()
            sel = gf != 0
            want[sel, seq.gripper_slot] = gf[sel]
            assert ((out["ctrl"] - want).abs() / scale).max().item() < 1e-6
            assert torch.equal(out["ctrl"][sel, seq.gripper_slot], gf[sel])
            recs.append({k: st[k][:n_chk].cpu().numpy().copy() for k in ("action", "err", "max_vel0", "target_xyz", "target_quat")})
        d = layout.as_dict()["devices"][ia]
        for i in range(n_chk):
            ps = {"active_xyz": poses[active][0][:, i], "active_quat": poses[active][1][:, i], "passive_xyz": poses[passive][0][:, i]}
            ref = sequence_numpy.run_sequence(seq.params, wp_xyz[i], wp_quat[i], ps, d, default_ee_quat(),
                                              layout.devices[ia].max_vel[0], T)
            for t in range(T):
                r, g = ref[t], recs[t]
                assert int(g["action"][i]) == r["action"], (active, i, t)
                assert g["max_vel0"][i] == pytest.approx(r["max_vel0"], rel=1e-12, abs=0)
                assert np.abs(g["target_xyz"][i, ia] - r["active_xyz"]).max() < 1e-15
                assert np.abs(g["target_xyz"][i, ip] - r["passive_xyz"]).max() < 1e-12
                assert np.abs(g["target_quat"][i, ip] - r["passive_quat"]).max() < 1e-15
            assert int(recs[-1]["action"][i]) == A
        # the unreachable episodes are still inside the first waypoint, at the saturated speed schedule
        a_all = st["action"].cpu().numpy()
        assert (a_all[n_chk:] == 0).all()


def test_waypoint_cycling_on_gpu_matches_the_gain_test_loop():
    import torch
    from irl_control_b200.engine import BatchedOSC
    from irl_control_b200.synthetic import scenario_model
    from oracle import sequence_numpy
    B, T, W, n_chk = 2048, 50, 4, 6
    dev = "cuda:0"
    layout, model = scenario_model("gain_test")
    names = [d.name for d in layout.devices]
    assert names == ["ur5right", "ur5left", "base"]
    eng = BatchedOSC(layout, device=0)
    eng.set_model(model)
    q, dq = _trajectory(B, T, seed=31)
    poses = _poses(layout, q[:, :n_chk])
    D = layout.D
    n_wp = [4, 3, 1]
    rng = np.random.default_rng(2)
    wps = rng.uniform(-0.5, 0.5, size=(B, D, W, 3)) + np.array([0.3, 0.0, 0.9])
    for d, ts in {0: [5, 11, 12, 30], 1: [3, 20, 33]}.items():
        for w, t in enumerate(ts):
            wps[:n_chk, d, w] = poses[names[d]][0][t]
    st = {"wps": torch.from_numpy(wps).to(dev), "n_wp": n_wp, "wp_idx": torch.zeros(B, D, dtype=torch.int32, device=dev),
          "target_xyz": torch.zeros(B, D, 3, dtype=torch.float64, device=dev),
          "target_quat": torch.tensor([1.0, 0, 0, 0], dtype=torch.float64, device=dev).expand(B, D, 4).contiguous()}
    mv = torch.tensor([list(d.max_vel) for d in layout.devices], dtype=torch.float64, device=dev)[None].expand(B, -1, -1).contiguous()
    qd, dqd = torch.from_numpy(q).to(dev), torch.from_numpy(dq).to(dev)
    got = []
    for t in range(T):
        before = st["wp_idx"][:n_chk].cpu().numpy().copy()
        out = eng.step_waypoints({"q": qd[t].contiguous(), "dq": dqd[t].contiguous(), "max_vel": mv}, st, threshold=0.1, want_u_all=True)
        ref = eng.step_fused({"q": qd[t].contiguous(), "dq": dqd[t].contiguous(), "max_vel": mv,
                              "target_xyz": st["target_xyz"], "target_quat": st["target_quat"]}, want_u_all=True)
        scale = ref["u_all"].abs().amax(dim=1, keepdim=True)
        assert ((out["u_all"] - ref["u_all"]).abs() / scale).max().item() < 1e-6      # two kernels: equal up to rounding
        got.append((st["target_xyz"][:n_chk].cpu().numpy().copy(), before))
    for i in range(n_chk):
        for d in (0, 1):
            ref = sequence_numpy.run_waypoint_cycle(wps[i, d, :n_wp[d]], poses[names[d]][0][:, i], 0.1, T)
            for t in range(T):
                assert got[t][1][i, d] == ref[t][1], (i, d, t)
                assert np.array_equal(got[t][0][i, d], ref[t][0])
