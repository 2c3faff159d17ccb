"""CPU: the host mirror of the reference API (index maps, error behaviour, layout flattening)."""
import os

import numpy as np
import pytest

import irl_control_b200 as pkg
from irl_control_b200.synthetic import SCENARIOS, build_scenario
from irl_control_b200.dual_ur5 import DualUR5Model, dynamics, sample_joint_states
from oracle.ref_harness import REFERENCE_ROOT as REF      # /root/reference, or its verbatim staged copy oracle/_ref


def _app(cfg="default_xyz_abg.yaml+start_body", scene="gain_test_scene.xml"):
    return pkg.MujocoApp(cfg, scene)


def test_device_index_maps_match_survey_a1():
    r = _app().get_robot("DualUR5")
    base, right, left = (r.get_device(n) for n in ("base", "ur5right", "ur5left"))
    assert list(base.joint_ids_all) == [0] and list(base.ctrl_idxs) == [0] and list(base.actuator_trnids) == [0]
    assert list(right.joint_ids) == [1, 2, 3, 4, 5, 6] and list(right.gripper_ids) == list(range(7, 13))
    assert list(right.ctrl_idxs) == [1, 2, 3, 4, 5, 6, 7] and list(right.actuator_trnids) == [1, 2, 3, 4, 5, 6, 10]
    assert list(left.joint_ids_all) == list(range(13, 25))
    assert list(left.ctrl_idxs) == list(range(8, 15)) and list(left.actuator_trnids) == [13, 14, 15, 16, 17, 18, 22]
    assert r.num_joints_total == 25 and list(r.joint_ids_all) == list(range(25))
    assert right.joint_names[0] == "joint0_ur5right"


def test_shipped_yaml_without_start_body_fails_like_reference():
    # 7 chain joints vs 6 start angles -> numpy shape error in Device.__init__
    with pytest.raises(ValueError):
        _app("default_xyz_abg.yaml")


def test_scene_free_objects_keep_robot_ids():
    app = _app(scene="insertion_task_scene.xml")
    assert app.sim.model.nv == 49
    r = app.get_robot("DualUR5")
    assert list(r.joint_ids_all) == list(range(25))
    M = r.get_state(pkg.RobotState.M)
    assert M.shape == (25, 25)
    Js, J_idxs = r.get_state(pkg.RobotState.J)
    assert Js["ur5right"].shape == (6, 25) and list(J_idxs["ur5left"]) == list(range(7, 13))


def test_layout_rows_and_dx_idx_order():
    for name, k, n_ctrl in (("gain_test", 7, 15), ("admit_test", 12, 14), ("insertion", 12, 14), ("worst_case", 13, 15),
                            ("iros2022", 13, 15)):
        _, _, targets, L = build_scenario(name)
        assert (L.k, L.n_ctrl, L.D) == (k, n_ctrl, len(targets))
    _, _, _, L = build_scenario("gain_test")
    # J_idxs follow sub-device order base, right, left (robot.py:52-55), targets are right, left, base
    assert [d.dx_idx for d in L.devices] == [(1, 2, 3), (4, 5, 6), (0,)]
    _, _, _, L = build_scenario("admit_test")
    # base is a sub-device of the robot even when it is not targeted -> left arm indexes past k = 12 (N3)
    assert L.admittance and [d.dx_idx for d in L.devices] == [tuple(range(1, 7)), tuple(range(7, 13))]


def test_osc_constructor_mutates_caller_config_like_reference():
    app = _app()
    cfg = app.get_controller_config("osc2")
    pkg.OSC(app.get_robot("DualUR5"), app.sim, [("ur5right", cfg)], app.get_controller_config("nullspace"))
    assert np.array_equal(cfg["task_space_gains"], [200] * 6) and np.allclose(cfg["lamb"], 4.0)


def test_thread_mode_asserts():
    app = pkg.MujocoApp("default_xyz_abg.yaml+start_body", "gain_test_scene.xml", use_sim=False)
    r = app.get_robot("DualUR5")
    osc = pkg.OSC(r, app.sim, [("ur5right", app.get_controller_config("osc2"))])
    with pytest.raises(AssertionError):
        osc.generate({"ur5right": pkg.Target()})      # osc.py:129-130
    app2 = _app()
    with pytest.raises(AssertionError):
        app2.get_robot("DualUR5").get_device("base").update_state()   # device.py:203
    with pytest.raises(AssertionError):
        app2.get_robot("DualUR5").stop()                              # robot.py:119


def test_unknown_names_raise_keyerror():
    app = _app()
    r = app.get_robot("DualUR5")
    with pytest.raises(KeyError):
        r.get_device("nope")
    osc = pkg.OSC(r, app.sim, [("ur5right", app.get_controller_config("osc2"))])
    with pytest.raises(KeyError):
        osc.layout_for(["ur5left"])     # no controller config for that device (osc.py:160)


def test_target_defaults_and_setters():
    t = pkg.Target()
    assert np.array_equal(t.get_quat(), [1, 0, 0, 0]) and np.array_equal(t.get_abg_vel(), [0, 0, 0])
    t.set_all_abg([1, 2, 3], [0.1, -0.2, 0.3])
    assert np.allclose(t.get_abg(), [0.1, -0.2, 0.3]) and np.array_equal(t.get_xyz(), [1, 2, 3])
    with pytest.raises(AssertionError):
        t.set_quat([1, 0, 0])
    assert t.velocity6().shape == (6,)


def test_calc_error_matches_oracle():
    from oracle import osc_numpy
    app, osc, targets, L = build_scenario("admit_test")
    sim = app.sim
    sim.data.qpos[:25] = sample_joint_states(1, 3)[0][0]
    sim.forward()
    dev = app.get_robot("DualUR5").get_device("ur5left")
    t = pkg.Target([0.1, 0.2, 0.3, 0.3, -0.4, 0.2])
    got = osc.calc_error(t, dev)
    want = osc_numpy.calc_error(L.as_dict()["devices"][1], dev.get_state(pkg.DeviceState.EE_XYZ),
                                dev.get_state(pkg.DeviceState.EE_QUAT), t.get_xyz(), t.get_quat())
    assert np.array_equal(got, want)


def test_dynamics_model_consistency():
    import torch
    m = DualUR5Model()
    q, dq = sample_joint_states(8, 5)
    q, dq = torch.from_numpy(q), torch.from_numpy(dq)
    d = dynamics(m, q, dq)
    assert torch.linalg.eigvalsh(d.M).min() > 0
    z = torch.zeros_like(dq)
    for j in (0, 3, 10, 24):
        e = torch.zeros_like(dq)
        e[:, j] = 1.0
        col = dynamics(m, q, z, ddq=e, need_M=False, gravity=(0, 0, 0)).bias
        assert (col - d.M[:, :, j]).abs().max() < 1e-12      # CRBA == RNEA columns
    ee = m.body_name2id("ur_EE_ur5left")
    jp, jr = d.jac_body(ee)
    assert jp[:, :, 1:13].abs().max() == 0 and jp[:, :, 19:].abs().max() == 0   # other arm / own gripper columns


def test_topology_is_derived_from_the_model_tree():
    """joint_parent / ee_joint come from the same body tree Device.__init__ walks (device.py:41-64)."""
    _, _, _, L = build_scenario("gain_test")
    assert L.joint_parent == (-1, 0, 1, 2, 3, 4, 5, 6, 7, 6, 6, 10, 6, 0, 13, 14, 15, 16, 17, 18, 19, 18, 18, 22, 18)
    assert [d.ee_joint for d in L.devices] == [6, 18, 0]
    p = L.to_c_params()
    assert p.has_topology == 1 and p.check_topology == 0 and p.dev[0].ee_joint == 6
    assert list(p.joint_parent)[:25] == list(L.joint_parent)
    # structural zeros promised by that tree really are exact zeros in the synthetic states
    import torch
    from irl_control_b200.synthetic import synth_batch
    st = synth_batch(L, 16, seed=2)
    anc = []
    for i in range(25):
        a, j = set(), i
        while j >= 0:
            a.add(j)
            j = L.joint_parent[j]
        anc.append(a)
    M = st["M"].numpy()
    for i in range(25):
        for j in range(i):
            if j not in anc[i]:
                assert (M[:, i, j] == 0).all()
    J = st["J6"].numpy()
    for d, dl in enumerate(L.devices):
        for j in range(25):
            if j not in anc[dl.ee_joint]:
                assert (J[:, d, :, j] == 0).all()


def test_bias_forces_satisfy_the_lagrangian_form():
    """Independent derivation of `qfrc_bias` (what osc.py:191 reads): the recursive Newton-Euler result of
    `dual_ur5.dynamics` must equal the Euler-Lagrange expression built from the inertia matrix and the potential
    energy alone, c_i = sum_jk (dM_ij/dq_k - 1/2 dM_jk/dq_i) dq_j dq_k + dV/dq_i, with the derivatives taken by
    central differences of M(q) and V(q) = -sum_b m_b g . com_b(q)."""
    import torch
    from irl_control_b200.dual_ur5 import GRAVITY
    m = DualUR5Model()
    qs, dqs = sample_joint_states(2, 9)
    n, h = 25, 1e-5
    g = np.asarray(GRAVITY)

    def M_and_V(qb):
        qt = torch.from_numpy(qb)
        d = dynamics(m, qt, torch.zeros_like(qt))
        V = np.zeros(qb.shape[0])
        for b in range(1, m.n_robot_bodies):
            it = m.body_inertial[b]
            if it is None:
                continue
            ipos, _iquat, mass, _diag = it
            com = d.xpos[:, b].numpy() + (d.xmat[:, b].numpy() @ np.asarray(ipos))
            V -= mass * (com @ g)
        return d.M.numpy(), V

    for q0, dq0 in zip(qs, dqs):
        qb = np.repeat(q0[None], 2 * n, 0)
        for k in range(n):
            qb[2 * k, k] += h
            qb[2 * k + 1, k] -= h
        Mb, Vb = M_and_V(qb)
        dM = (Mb[0::2] - Mb[1::2]) / (2 * h)                      # dM[k][i][j] = dM_ij / dq_k
        dV = (Vb[0::2] - Vb[1::2]) / (2 * h)
        c = np.einsum("kij,j,k->i", dM, dq0, dq0) - 0.5 * np.einsum("ijk,j,k->i", dM, dq0, dq0) + dV
        bias = dynamics(m, torch.from_numpy(q0[None]), torch.from_numpy(dq0[None])).bias.numpy()[0]
        assert np.abs(bias - c).max() < 1e-6 * max(1.0, np.abs(bias).max()), np.abs(bias - c).max()


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "irl_control/scenes/dual_ur5.xml")), reason="needs /root/reference")
def test_model_table_is_what_the_extractor_reads_from_the_reference_scene():
    """`dual_ur5_model.py` (travels with the repo) == a fresh extraction from the reference's dual_ur5.xml."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "extract_dual_ur5.py"),
                          os.path.join(REF, "irl_control/scenes/dual_ur5.xml")], capture_output=True, text=True, check=True).stdout
    with open(os.path.join(root, "irl_control_b200", "dual_ur5_model.py")) as fh:
        assert out.strip() == fh.read().strip()


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "irl_control/robot_configs")), reason="needs /root/reference")
def test_builtin_configs_equal_the_reference_yaml_files():
    """The Python copies in configs.py (they travel to the GPU box) carry the reference's YAML values."""
    import yaml
    from irl_control_b200 import configs
    ref = os.path.join(REF, "irl_control")
    for name in ("default_xyz.yaml", "default_xyz_abg.yaml", "iros2022.yaml"):
        with open(os.path.join(ref, "robot_configs", name)) as fh:
            want = yaml.safe_load(fh)
        assert configs.robot_config(name) == want, name
    with open(os.path.join(ref, "action_sequence_configs", "insertion_task.yaml")) as fh:
        assert configs.action_config("insertion_task.yaml") == yaml.safe_load(fh)
    with open(os.path.join(ref, "action_sequence_configs", "iros2022_task.yaml")) as fh:
        assert configs.IROS2022_DEVICE_CONFIG == yaml.safe_load(fh)["device_config"]


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "irl_control/osc.py")), reason="needs /root/reference")
@pytest.mark.parametrize("scenario", sorted(SCENARIOS))
def test_host_classes_resolve_the_same_maps_as_the_reference_classes(scenario):
    """`Device` / `Robot` / `OSC` of this package next to the reference's own classes constructed on the same model
    and YAML: index maps, DoF masks, Jacobian row numbering, gains after the constructor's precompute."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    runner = make_golden.reference_runner(scenario)
    app, osc, targets, layout = build_scenario(scenario)
    robot = app.get_robot("DualUR5")
    assert robot.num_joints_total == runner.robot.num_joints_total
    assert list(robot.joint_ids_all) == list(runner.robot.joint_ids_all)
    for name, ref_dev in runner.robot.sub_devices_dict.items():
        dev = robot.get_device(name)
        for attr in ("joint_ids", "joint_ids_all", "ctrl_idxs", "actuator_trnids", "ctrlr_dof", "ctrlr_dof_xyz",
                     "ctrlr_dof_abg", "max_vel", "start_angles"):
            assert list(np.ravel(getattr(dev, attr))) == list(np.ravel(getattr(ref_dev, attr))), (name, attr)
    # Jacobian row numbering (robot.py:52-55) as seen through the layout's dx_idx
    from irl_control_b200.synthetic import synth_batch
    st = {k: v.numpy()[0] for k, v in synth_batch(layout, 1, seed=3).items()}
    runner.run(st, st["target_xyz"], st["target_quat"], max_vel=st["max_vel"])          # loads the instance
    _Js, J_idxs = runner.robot.get_all_states()[make_golden.ref_harness.import_reference()[5].J]
    for dl in layout.devices:
        assert list(dl.dx_idx) == list(J_idxs[dl.name]), dl.name
    # gains after OSC.__init__'s in-place precompute (osc.py:35-39)
    for dl in layout.devices:
        cfg = runner.osc.device_configs[dl.name]
        assert (dl.kp, dl.kv, dl.ko) == (cfg["kp"], cfg["kv"], cfg["ko"])
        assert list(dl.k) == list(cfg["k"]) and list(dl.d) == list(cfg["d"])
    assert layout.nullspace_kv == runner.osc.nullspace_config["kv"]


def test_roofline_bytes_are_the_survey_figures():
    """`bench.algorithmic_bytes` (the numerator of roofline.achieved) against SURVEY.md 8(d): 4 904 B gain_test,
    5 864 B admit_test, 5 768 B insertion, 6 104 B for k = 13 (worst case and iros2022)."""
    import bench
    want = {"gain_test": 4904, "admit_test": 5864, "insertion": 5768, "worst_case": 6104, "iros2022": 6104}
    for name, nbytes in want.items():
        _, _, _, L = build_scenario(name)
        assert bench.algorithmic_bytes(L) == nbytes, name


def test_mujoco_app_helpers():
    """`get_robot`, `get_controller_config`, `set_free_joint_qpos`, `sleep_for` (mujoco_app.py:37-63)."""
    import threading
    import time
    app = pkg.MujocoApp("default_xyz_abg.yaml+start_body", "insertion_task_scene.xml")
    assert app.get_robot("DualUR5").name == "DualUR5" and app.get_robot("nope") is None
    cfg = app.get_controller_config("osc1")
    assert cfg["kv"] == 50 and cfg is app.get_controller_config("osc1") and app.get_controller_config("zzz") is None
    m = app.sim.model
    name = m.joint_id2name(25)                                    # first free joint after the robot's 25 hinges
    off = m.jnt_qposadr[m.joint_name2id(name)]
    app.set_free_joint_qpos(name, quat=[0.5, 0.5, 0.5, 0.5], pos=[1, 2, 3])
    assert list(app.sim.data.qpos[off:off + 7]) == [1, 2, 3, 0.5, 0.5, 0.5, 0.5]
    app.set_free_joint_qpos(name, pos=[4, 5, 6])
    assert list(app.sim.data.qpos[off:off + 7]) == [4, 5, 6, 0.5, 0.5, 0.5, 0.5]
    th = threading.Thread(target=app.sleep_for, args=(0.2,))
    th.start()
    time.sleep(0.05)
    assert app.timer_running
    with pytest.raises(AssertionError):                           # not re-entrant (mujoco_app.py:38)
        app.sleep_for(0.01)
    th.join()
    assert not app.timer_running
