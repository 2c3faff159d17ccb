"""GPU: BASELINE config 4 - the reference's 12-entry insertion action list (insertion_task.yaml:35-104) with
adapters placed at random per episode (insertion_task.py:341-369), B = 16 384 episodes advanced by
`irlosc_step_sequence`.

Written after round 1's GPU budget was spent; the same list runs through the host build of the kernel in
tests/test_insertion_host.py.  Collected last because this is its first run on a GPU.
"""
import numpy as np
import pytest

from test_sequence_host import _poses, _trajectory

pytestmark = pytest.mark.gpu


def test_reference_action_list_batch_16384():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from irl_control_b200 import insertion
    from irl_control_b200.configs import action_config
    from irl_control_b200.engine import BatchedOSC
    from irl_control_b200.sequence import ActionSequence, default_ee_quat
    from irl_control_b200.synthetic import scenario_model
    from oracle import sequence_numpy
    B, T, n_chk = 16384, 110, 6
    dev = "cuda:0"
    cfg = action_config("insertion_task.yaml")
    actions, objs = cfg["insertion_action_sequence"], cfg["nist_action_objects"]
    A = len(actions)
    layout, model = scenario_model("insertion")
    names = [d.name for d in layout.devices]
    active, passive = "ur5right", "ur5left"
    ia, ip = names.index(active), names.index(passive)
    eng = BatchedOSC(layout, device=0)
    eng.set_model(model)
    seq = ActionSequence(layout, actions, active_arm=active, step_period=0.25)
    q, dq = _trajectory(B, T, seed=12)
    poses = _poses(layout, q[:, :n_chk])
    # episodes >= n_chk: waypoints from randomly placed adapters (most are not reached within T steps);
    # episodes < n_chk: waypoints on the prescribed trajectory so that all twelve actions complete and are checked
    placed = insertion.random_object_poses(B, "right", objs, rng=np.random.default_rng(3))
    start = np.zeros((B, 3))
    start[:n_chk] = poses[active][0][0]
    wp_xyz, wp_quat = insertion.waypoint_poses(actions, objs, placed, start)
    wp_actions = [a for a, p in enumerate(actions) if p["action"] == "WP"]
    for a, t in zip(wp_actions, (5, 16, 31, 38, 46, 58, 74, 95)):
        wp_xyz[:n_chk, a], wp_quat[:n_chk, a] = poses[active][0][t], poses[active][1][t]
    st = seq.new_state(B, wp_xyz, wp_quat, device=dev)
    mv = torch.tensor([list(d.max_vel) for d in layout.devices], dtype=torch.float64, device=dev)[None].expand(B, -1, -1).contiguous()
    qd, dqd = torch.from_numpy(q).to(dev), torch.from_numpy(dq).to(dev)
    recs, grip = [], []
    for t in range(T):
        out = eng.step_sequence({"q": qd[t].contiguous(), "dq": dqd[t].contiguous(), "max_vel": mv}, seq, st)
        assert torch.isfinite(out["ctrl"]).all()
        recs.append({k: st[k][:n_chk].cpu().numpy().copy() for k in ("action", "err", "max_vel0", "target_xyz", "target_quat")})
        grip.append(out["ctrl"][:n_chk, seq.gripper_slot].cpu().numpy().copy())
    d = layout.as_dict()["devices"][ia]
    for i in range(n_chk):
        ps = {"active_xyz": poses[active][0][:, i], "active_quat": poses[active][1][:, i], "passive_xyz": poses[passive][0][:, i]}
        ref = sequence_numpy.run_sequence(seq.params, wp_xyz[i], wp_quat[i], ps, d, default_ee_quat(),
                                          layout.devices[ia].max_vel[0], T)
        for t in range(T):
            r = ref[t]
            assert int(recs[t]["action"][i]) == r["action"], (i, t)
            assert recs[t]["max_vel0"][i] == pytest.approx(r["max_vel0"], rel=1e-9, abs=0)
            assert np.abs(recs[t]["target_xyz"][i][ia] - r["active_xyz"]).max() < 1e-15
            assert np.abs(recs[t]["target_quat"][i][ia] - r["active_quat"]).max() < 1e-15
            assert np.abs(recs[t]["target_xyz"][i][ip] - r["passive_xyz"]).max() < 1e-9
            if r["gripper_force"] != 0.0:
                assert grip[t][i] == r["gripper_force"]
        assert int(recs[-1]["action"][i]) == A
    acts = st["action"].cpu().numpy()
    assert acts.min() >= 0 and acts.max() <= A
