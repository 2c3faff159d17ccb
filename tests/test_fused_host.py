"""CPU: arithmetic of the fused state provider + OSC step.

`irl_control_b200/csrc/osc_fused.cuh` keeps its per-instance function `__host__ __device__`;
tests/host_fused compiles exactly that function for the host (test infrastructure, never loaded
by the package) so this suite can compare it with the oracle without a GPU:

  * the provider part - M dq, qfrc_bias, J rows, dx = J dq, A = J M^-1 J^T, EE poses - against the
    rigid-body model in `dual_ur5.py` (what the reference reads from MuJoCo);
  * the joint-space signal against `oracle/osc_numpy.py` fed with that model's M / J / bias.

The GPU suite (tests/test_gpu_fused.py) then checks the compiled kernel against the same oracle.
"""
import numpy as np
import pytest

import fused_host
from irl_control_b200.dual_ur5 import sample_joint_states
from irl_control_b200.rigid_model import model_for_layout
from irl_control_b200 import _native
from irl_control_b200.synthetic import build_scenario, oracle_inputs, scenario_layout, synth_batch
from oracle import osc_numpy

REL_TOL = 1e-6      # same bar as tests/test_gpu_parity.py (BASELINE.json asks 1e-4)


def _case(scenario, B, seed, with_vel=None):
    app, _osc, _names, layout = build_scenario(scenario)
    robot = app.get_robot("DualUR5")
    model = model_for_layout(app.sim.model, robot.joint_ids_all, layout)
    st = synth_batch(layout, B, seed=seed, insertion_schedule=(scenario == "insertion"))
    q, dq = sample_joint_states(B, seed)
    assert np.array_equal(q, st["q"].numpy()) and np.array_equal(dq, st["dq"].numpy())
    inp = {"q": q, "dq": dq, "target_xyz": st["target_xyz"].numpy(), "target_quat": st["target_quat"].numpy(),
           "max_vel": st["max_vel"].numpy()}
    if layout.admittance:
        inp["ft_raw"] = st["ft_raw"].numpy()
    if with_vel is not None:
        import torch
        st["target_vel"] = torch.from_numpy(with_vel)
        inp["target_vel"] = with_vel
    return layout, model, st, inp


def _rel(got, want):
    scale = np.abs(want).max(axis=1, keepdims=True)
    return (np.abs(got - want) / scale).max(axis=1)


@pytest.mark.parametrize("scenario", ["gain_test", "admit_test", "insertion", "worst_case", "iros2022"])
def test_provider_quantities_match_the_rigid_body_model(scenario):
    layout, model, st, inp = _case(scenario, 96, seed=11)
    out = fused_host.run(layout, model, inp, debug=True)
    M, J, dq = st["M"].numpy(), st["J"].numpy(), st["dq"].numpy()
    assert np.abs(out["uv"] - np.einsum("bij,bj->bi", M, dq)).max() < 1e-12
    assert np.abs(out["bias"] - st["bias"].numpy()).max() < 1e-11
    assert np.abs(out["J"] - J).max() < 1e-13
    assert np.abs(out["dx"] - np.einsum("bkj,bj->bk", J, dq)).max() < 1e-13
    A = J @ np.linalg.solve(M, J.transpose(0, 2, 1))
    assert (np.abs(out["A"] - A).max(axis=(1, 2)) / np.abs(A).max(axis=(1, 2))).max() < 1e-11
    assert np.abs(out["ee_xyz"] - st["ee_xyz"].numpy()).max() < 1e-13
    eq = st["ee_quat"].numpy()
    sgn = np.sign((out["ee_quat"] * eq).sum(-1, keepdims=True))
    assert np.abs(out["ee_quat"] * sgn - eq).max() < 1e-13


@pytest.mark.parametrize("scenario,B", [("gain_test", 512), ("admit_test", 512), ("insertion", 512), ("worst_case", 256),
                                        ("iros2022", 256)])
def test_fused_step_matches_the_oracle(scenario, B):
    layout, model, st, inp = _case(scenario, B, seed=3)
    out = fused_host.run(layout, model, inp)
    ref = osc_numpy.osc_batch(layout.as_dict(), oracle_inputs(st, layout))
    rel = _rel(out["u_all"], ref["u_all"])
    # instances within rounding of the pinv cutoff may legitimately differ; none are expected here
    assert rel.max() < REL_TOL, (rel.max(), int(np.argmax(rel)))
    assert np.array_equal((out["status"] & 1).astype(bool), np.asarray(ref["pinv"]).astype(bool))
    # packing (osc.py:203-208)
    want = np.concatenate([ref["u_all"][:, list(d.actuator_trnids)] for d in layout.devices], axis=1)
    assert _rel(out["ctrl"], want).max() < REL_TOL
    if scenario in ("admit_test", "insertion", "worst_case", "iros2022"):
        assert (out["how"] & (fused_host.HOW_CUT1 | fused_host.HOW_CUT2)).any()   # the pinv deflation is exercised


def test_velocity_tracking_branch_and_index_error_flag():
    """N4: the branch is taken only when all six target-velocity entries are non-zero; N3: in the
    gain_test target order J_idxs runs past k for the left arm -> IndexError in the reference."""
    from irl_control_b200 import _native
    B = 64
    rng = np.random.default_rng(0)
    for scenario in ("admit_test", "gain_test"):
        app, _o, _n, layout = build_scenario(scenario)
        tv = rng.normal(0.0, 0.2, size=(B, layout.D, 6))
        tv[: B // 2, :, 0] = 0.0                      # one zero entry -> "zero" branch (osc.py:173)
        layout, model, st, inp = _case(scenario, B, seed=9, with_vel=tv)
        out = fused_host.run(layout, model, inp)
        ob = oracle_inputs(st, layout)
        n_err = n_track = 0
        for i in range(B):
            try:
                ref = osc_numpy.osc_step(layout.as_dict(), {k: v[i] for k, v in ob.items()})
            except IndexError:                           # robot.py:52-55 vs osc.py:150,176
                n_err += 1
                assert out["status"][i] & _native.ST_DX_RANGE
                assert np.isnan(out["u_all"][i]).all() and np.isnan(out["ctrl"][i]).all()
                continue
            assert not out["status"][i] & _native.ST_DX_RANGE
            assert _rel(out["u_all"][i:i + 1], ref["u_all"][None]).max() < REL_TOL
            tracking = bool(np.any(ref["vel_branch"]))
            n_track += tracking
            assert bool(out["status"][i] & _native.ST_VEL_BRANCH) == tracking
        assert n_err + n_track == B // 2, (scenario, n_err, n_track)


def test_reduced_model_lumps_welded_bodies():
    from irl_control_b200.dual_ur5 import DualUR5Model
    app, _o, _n, layout = build_scenario("gain_test")
    robot = app.get_robot("DualUR5")
    m = app.sim.model
    model = model_for_layout(m, robot.joint_ids_all, layout)
    assert model.n_joints == 25
    assert [model.joint[j].parent for j in range(25)] == [-1, 0, 1, 2, 3, 4, 5, 6, 7, 6, 6, 10, 6,
                                                          0, 13, 14, 15, 16, 17, 18, 19, 18, 18, 22, 18]
    moving = sum(it[2] for b, it in enumerate(m.body_inertial) if it is not None and b < m.n_robot_bodies) - 100.0
    assert abs(sum(model.joint[j].mass for j in range(25)) - moving) < 1e-12     # origin_base (100 kg) is welded to the world
    assert [model.ee[d].joint for d in range(3)] == [6, 18, 0]
    assert [model.ft[d].joint for d in range(3)] == [6, 18, -1]


def test_mujoco_binding_view_reduces_to_the_same_model():
    """`mujoco_adapter.MjModelView` on a stand-in that carries the arrays of `mujoco.MjModel`
    (mujoco itself is absent): the reduced model equals the one built from the mujoco_py-shaped model."""
    import ctypes as C
    from types import SimpleNamespace
    from irl_control_b200.mujoco_adapter import MjModelView, joint_state, scatter_ctrl
    from irl_control_b200.rigid_model import reduce_model
    app, _o, names, layout = build_scenario("admit_test")
    robot = app.get_robot("DualUR5")
    m = app.sim.model
    nb = m.n_robot_bodies
    mass, ipos, iquat, inertia = np.zeros(nb), np.zeros((nb, 3)), np.tile([1.0, 0, 0, 0], (nb, 1)), np.zeros((nb, 3))
    for b, it in enumerate(m.body_inertial[:nb]):
        if it is not None:
            ipos[b], iquat[b], mass[b], inertia[b] = it
    mj = SimpleNamespace(nbody=nb, body_parentid=m.body_parentid[:nb], body_pos=m.body_pos[:nb], body_quat=m.body_quat[:nb],
                         body_jntadr=m.body_jntadr[:nb], body_jntnum=m.body_jntnum[:nb], jnt_bodyid=m.jnt_bodyid,
                         jnt_axis=m.jnt_axis, jnt_pos=m.jnt_pos, jnt_qposadr=np.arange(25), jnt_dofadr=np.arange(25),
                         site_bodyid=m.site_bodyid, site_pos=m.site_pos, site_quat=m.site_quat,
                         body_mass=mass, body_ipos=ipos, body_iquat=iquat, body_inertia=inertia)
    ids = {"body": m.body_name2id, "site": m.site_name2id}
    view = MjModelView(mj, name2id=lambda kind, name: ids[kind](name))
    ee = [d.ee_body for d in layout.devices]
    ft = ["ft_frame_" + d.name for d in layout.devices]
    a = reduce_model(view, robot.joint_ids_all, ee, ft)
    b = model_for_layout(m, robot.joint_ids_all, layout)
    assert bytes(a) == bytes(b)
    assert "ft_frame_ur5right" in view.site_names and "nope" not in view.site_names
    data = SimpleNamespace(qpos=np.arange(40.0), qvel=-np.arange(37.0), ctrl=np.zeros(15))
    q, dq = joint_state(data, view, robot.joint_ids_all)
    assert np.array_equal(q, np.arange(25.0)) and np.array_equal(dq, -np.arange(25.0))
    row = np.arange(1.0, layout.n_ctrl + 1)
    scatter_ctrl(data.ctrl, layout, row)
    for sl, dl in zip(layout.ctrl_slices, layout.devices):
        assert np.array_equal(data.ctrl[list(dl.ctrl_idxs)], row[sl])


@pytest.mark.parametrize("scenario,B,min_cut1,max_warp", [("gain_test", 2048, 0, 0), ("admit_test", 4096, 30, 0),
                                                          ("worst_case", 4096, 300, 1)])
def test_task_space_solve_is_decided_in_the_thread(scenario, B, min_cut1, max_warp):
    """osc_tail.cuh resolves osc.py:52-55 on the block structure of A in the thread that owns the instance:
    exact inertia counts decide how many eigenvalues numpy's pinv(rcond=1e-5) cuts, deflation removes them.
    Against numpy on the same A: every decided instance cuts exactly the eigenvalues numpy cuts, the result
    matches the oracle, and only a handful are left to the warp-cooperative eigen-solver."""
    layout = scenario_layout(scenario)
    st = synth_batch(layout, B, seed=29)
    state = {k: st[k].numpy() for k in ("M", "J", "dq", "bias", "ee_xyz", "ee_quat", "target_xyz", "target_quat", "max_vel")}
    if layout.admittance:
        state["ft_xmat"], state["ft_raw"] = st["ft_xmat"].numpy(), st["ft_raw"].numpy()
    out = fused_host.run_stream(layout, state)
    ref = osc_numpy.osc_batch(layout.as_dict(), oracle_inputs(st, layout))
    assert _rel(out["u_all"], ref["u_all"]).max() < REL_TOL
    M, J = state["M"], state["J"]
    A = J @ np.linalg.solve(M, J.transpose(0, 2, 1))
    ev = np.linalg.eigvalsh(0.5 * (A + A.transpose(0, 2, 1)))
    small = np.abs(np.linalg.det(A)) < 1e-4
    n_cut = np.where(small, (ev <= 1e-5 * ev[:, -1:]).sum(1), 0)
    how = out["how"]
    decided = how != fused_host.HOW_WARP
    want = np.select([n_cut == 0, n_cut == 1, n_cut == 2], [fused_host.HOW_INVERSE, fused_host.HOW_CUT1, fused_host.HOW_CUT2], -1)
    assert np.array_equal(how[decided], want[decided])
    assert (how == fused_host.HOW_CUT1).sum() >= min_cut1
    assert (~decided).sum() <= max_warp and out["n_hard"] == (~decided).sum()
    assert np.array_equal((out["status"] & _native.ST_PINV) != 0, small)


@pytest.mark.parametrize("scenario,use_g,nullspace,no_max_vel", [
    ("gain_test", False, True, ()), ("admit_test", True, False, ()), ("worst_case", False, False, ()),
    ("gain_test", True, True, ("ur5left", "base")), ("insertion", True, True, ("ur5right",))])
def test_fused_step_with_constructor_options_and_cleared_max_vel(scenario, use_g, nullspace, no_max_vel):
    """The fused step under the options no example uses - `use_g=False` (osc.py:190), `nullspace_config=None`
    (osc.py:195), `device.max_vel = None` (osc.py:163-168) - against the oracle, whose handling of exactly these
    options is pinned by the reference goldens `*_no_g_*`, `*_no_nullspace_*`, `*_bare_*`, `*_nomaxvel_*`."""
    import dataclasses
    B = 128
    app, _osc, _names, layout = build_scenario(scenario)
    layout = dataclasses.replace(
        layout, use_g=use_g, nullspace_kv=layout.nullspace_kv if nullspace else None,
        devices=tuple(dataclasses.replace(d, has_max_vel=d.name not in no_max_vel) for d in layout.devices))
    robot = app.get_robot("DualUR5")
    model = model_for_layout(app.sim.model, robot.joint_ids_all, layout)
    st = synth_batch(layout, B, seed=17, insertion_schedule=(scenario == "insertion"))
    inp = {"q": st["q"].numpy(), "dq": st["dq"].numpy(), "target_xyz": st["target_xyz"].numpy(),
           "target_quat": st["target_quat"].numpy(), "max_vel": st["max_vel"].numpy()}
    if layout.admittance:
        inp["ft_raw"] = st["ft_raw"].numpy()
    out = fused_host.run(layout, model, inp)
    ref = osc_numpy.osc_batch(layout.as_dict(), oracle_inputs(st, layout))
    assert _rel(out["u_all"], ref["u_all"]).max() < REL_TOL
    assert np.array_equal((out["status"] & 1).astype(bool), np.asarray(ref["pinv"]).astype(bool))
