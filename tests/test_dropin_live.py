"""CPU, build container only: the drop-in boundary exercised next to the live reference.

The reference's own `Device / Robot / OSC` and this package's classes are constructed on the SAME fake simulator
(oracle/ref_harness.FakeSim, the mujoco_py-shaped accessors of SURVEY 8c) from the SAME reference YAML, the same
instance is loaded, and `generate(targets)` is called on both.  Everything above the C ABI is the product's code
(`Device.get_state`, `Robot.get_all_states`, `OSC.gather_state`, packing, error behaviour); below it the stand-in
engine runs the host build of the streaming step (tests/host_fused) because there is no GPU here - the CUDA build
of the same function is what tests/test_gpu_parity.py checks.
"""
import os
import sys

import numpy as np
import pytest

import fused_host
import irl_control_b200 as pkg
import irl_control_b200.osc as pkg_osc
from irl_control_b200.synthetic import SCENARIOS, build_scenario, synth_batch
from oracle import ref_harness

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_harness.reference_available(), reason="needs /root/reference")]


class _HostEngine:
    def __init__(self, layout, device=None):
        self.layout, self.k, self.n_ctrl = layout, layout.k, layout.n_ctrl

    def step_host(self, state, **kw):
        st = dict(state)
        if "ft_xmat" in st:
            st["ft_xmat"] = st["ft_xmat"].reshape(st["ft_xmat"].shape[0], -1, 9)
        return fused_host.run_stream(self.layout, st)


# mixed_dof (generic kernel) and worst_case_admit (copy plan too large for the streaming kernel) have no host build: GPU only
@pytest.mark.parametrize("scenario", sorted(set(SCENARIOS) - {"mixed_dof", "worst_case_admit"}))
def test_generate_equals_the_reference_generate_on_the_same_simulator(scenario, monkeypatch):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    monkeypatch.setattr(pkg_osc, "BatchedOSC", _HostEngine)
    sc = SCENARIOS[scenario]
    runner = make_golden.reference_runner(scenario)
    sim, model = runner.sim, runner.sim.model
    cfg = ref_harness.load_reference_yaml(sc["config"].replace("+start_body", ""), inject_start_body=True)
    devices = [pkg.Device(d, model, sim, True) for d in cfg["devices"]]
    robot = pkg.Robot([devices[i] for i in cfg["robots"][0]["device_ids"]], "DualUR5", sim, True)
    by_name = {c["name"]: c for c in cfg["controller_configs"]}
    osc = pkg.OSC(robot, sim, [(dev, dict(by_name[c])) for dev, c in sc["device_cfgs"]], dict(by_name["nullspace"]),
                  admittance=sc["admittance"])
    _, _, names, layout = build_scenario(scenario)
    B = 6
    st = {k: v.numpy() for k, v in synth_batch(layout, B, seed=41, insertion_schedule=(scenario == "insertion")).items()}
    worst = 0.0
    for i in range(B):
        inst = {k: v[i] for k, v in st.items()}
        r = runner.run(inst, st["target_xyz"][i], st["target_quat"][i], max_vel=st["max_vel"][i])     # loads the sim
        targets = {}
        for d, nm in enumerate(names):
            t = pkg.Target()
            t.set_xyz(st["target_xyz"][i][d])
            t.set_quat(st["target_quat"][i][d])
            targets[nm] = t
            robot.get_device(nm).max_vel = [float(st["max_vel"][i][d][0]), float(st["max_vel"][i][d][1])]
        idxs, forces = osc.generate(targets)
        assert len(forces) == len(r["forces"]) == len(names)
        scale = np.abs(r["u_all"]).max()
        for d in range(len(names)):
            assert list(idxs[d]) == list(r["ctrl_idxs"][d])
            worst = max(worst, np.abs(forces[d] - r["forces"][d]).max() / scale)
    assert worst < 1e-6, worst


def _both(scenario, monkeypatch):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    monkeypatch.setattr(pkg_osc, "BatchedOSC", _HostEngine)
    sc = SCENARIOS[scenario]
    runner = make_golden.reference_runner(scenario)
    sim, model = runner.sim, runner.sim.model
    cfg = ref_harness.load_reference_yaml(sc["config"].replace("+start_body", ""), inject_start_body=True)
    devices = [pkg.Device(d, model, sim, True) for d in cfg["devices"]]
    robot = pkg.Robot([devices[i] for i in cfg["robots"][0]["device_ids"]], "DualUR5", sim, True)
    by_name = {c["name"]: c for c in cfg["controller_configs"]}
    osc = pkg.OSC(robot, sim, [(dev, dict(by_name[c])) for dev, c in sc["device_cfgs"]], dict(by_name["nullspace"]),
                  admittance=sc["admittance"])
    _, _, names, layout = build_scenario(scenario)
    return runner, osc, names, layout


def test_target_velocity_branch_and_index_error_like_the_reference(monkeypatch):
    """Non-zero target velocities through both `generate`s: the tracking branch with the reference's J_idxs
    numbering (N3 / N4) on gain_test, and the IndexError the reference raises on the insertion layout."""
    ref_Target = ref_harness.import_reference()[3]
    runner, osc, names, layout = _both("gain_test", monkeypatch)
    st = {k: v.numpy() for k, v in synth_batch(layout, 3, seed=8).items()}
    rng = np.random.default_rng(3)
    for i in range(3):
        runner.run({k: v[i] for k, v in st.items()}, st["target_xyz"][i], st["target_quat"][i], max_vel=st["max_vel"][i])
        vel = rng.uniform(0.05, 0.3, size=(len(names), 6)) * rng.choice([-1.0, 1.0], size=(len(names), 6))
        if i == 2:
            vel[0, 4] = 0.0                                   # one zero component: that device keeps the zero branch
        mine, theirs = {}, {}
        for d, nm in enumerate(names):
            for cls, bag in ((pkg.Target, mine), (ref_Target, theirs)):
                t = cls(np.zeros(6), vel[d])
                t.set_xyz(st["target_xyz"][i][d])
                t.set_quat(st["target_quat"][i][d])
                bag[nm] = t
        ridx, rforces = runner.osc.generate(theirs)
        idxs, forces = osc.generate(mine)
        scale = max(np.abs(f).max() for f in rforces)
        for d in range(len(names)):
            assert list(idxs[d]) == list(ridx[d])
            assert np.abs(forces[d] - rforces[d]).max() < 1e-6 * scale
    runner, osc, names, layout = _both("insertion", monkeypatch)
    st = {k: v.numpy() for k, v in synth_batch(layout, 1, seed=9).items()}
    runner.run({k: v[0] for k, v in st.items()}, st["target_xyz"][0], st["target_quat"][0], max_vel=st["max_vel"][0])
    vel = np.full(6, 0.1)
    with pytest.raises(IndexError):
        runner.osc.generate({nm: ref_Target(np.zeros(6), vel) for nm in names})
    with pytest.raises(IndexError):
        osc.generate({nm: pkg.Target(np.zeros(6), vel) for nm in names})


def test_max_vel_none_takes_the_unsaturated_gain_branch_like_the_reference(monkeypatch):
    """`device.max_vel = None` (a caller may clear it; osc.py:163-168): the pose error is multiplied by the task-space
    gains and the stiffness without the velocity limiter - also after a generate() with the limiter has already run."""
    runner, osc, names, layout = _both("gain_test", monkeypatch)
    ref_Target = ref_harness.import_reference()[3]
    st = {k: v.numpy() for k, v in synth_batch(layout, 2, seed=15).items()}
    robot = osc.robot
    for i, cleared in ((0, ()), (1, ("ur5left",)), (1, ("ur5left", "base"))):
        runner.run({k: v[i] for k, v in st.items()}, st["target_xyz"][i], st["target_quat"][i], max_vel=st["max_vel"][i])
        mine, theirs = {}, {}
        for d, nm in enumerate(names):
            for cls, bag in ((pkg.Target, mine), (ref_Target, theirs)):
                t = cls()
                t.set_xyz(st["target_xyz"][i][d])
                t.set_quat(st["target_quat"][i][d])
                bag[nm] = t
            mv = None if nm in cleared else [float(st["max_vel"][i][d][0]), float(st["max_vel"][i][d][1])]
            robot.get_device(nm).max_vel = mv
            runner.robot.get_device(nm).max_vel = mv
        ridx, rforces = runner.osc.generate(theirs)
        idxs, forces = osc.generate(mine)
        scale = max(np.abs(f).max() for f in rforces)
        for d in range(len(names)):
            assert np.abs(forces[d] - rforces[d]).max() < 1e-6 * scale, (cleared, names[d])


def test_gains_edited_after_construction_behave_like_the_reference(monkeypatch):
    """`OSC.__init__` stores `task_space_gains` and `lamb` in the config dict once (osc.py:35-39); `kp / kv / ko` are read
    fresh every step (osc.py:76,170).  A caller that edits a gain afterwards therefore gets the NEW saturation / kv
    factors with the OLD lamb - in the reference and here (also when a device has no velocity limit: old gains)."""
    runner, osc, names, layout = _both("gain_test", monkeypatch)
    ref_Target = ref_harness.import_reference()[3]
    st = {k: v.numpy() for k, v in synth_batch(layout, 2, seed=23).items()}
    for step, (edit, cleared) in enumerate(((dict(kp=320.0, kv=35.0), ()), (dict(ko=90.0, kv=12.0), ("ur5right",)))):
        for o in (osc, runner.osc):
            for nm in ("ur5right", "ur5left"):
                for key, val in edit.items():
                    o.device_configs[nm][key] = val
        i = step
        runner.run({k: v[i] for k, v in st.items()}, st["target_xyz"][i], st["target_quat"][i], max_vel=st["max_vel"][i])
        mine, theirs = {}, {}
        for d, nm in enumerate(names):
            for cls, bag in ((pkg.Target, mine), (ref_Target, theirs)):
                t = cls()
                t.set_xyz(st["target_xyz"][i][d])
                t.set_quat(st["target_quat"][i][d])
                bag[nm] = t
            mv = None if nm in cleared else [float(st["max_vel"][i][d][0]), float(st["max_vel"][i][d][1])]
            osc.robot.get_device(nm).max_vel = mv
            runner.robot.get_device(nm).max_vel = mv
        ridx, rforces = runner.osc.generate(theirs)
        idxs, forces = osc.generate(mine)
        scale = max(np.abs(f).max() for f in rforces)
        for d in range(len(names)):
            assert np.abs(forces[d] - rforces[d]).max() < 1e-6 * scale, (step, names[d])
    # the stored vectors are the ones of construction time, not of the edited gains
    assert np.allclose(osc.device_configs["ur5right"]["lamb"], np.array([200.0] * 6) / 20.0)
