"""CPU: the drop-in boundary against the reference's golden outputs, without the reference.

This package's `Device / Robot / OSC` are constructed on a fake mujoco_py-shaped simulator (oracle/ref_harness.FakeSim,
pure numpy) from the built-in copies of the reference YAMLs; every golden instance is loaded into the simulator and
`OSC.generate(targets)` must return the forces the unmodified reference returned for it (tests/golden/*.npz).  Above
the C ABI everything is the product's code; the stand-in engine below it runs the host build of the streaming step
(tests/host_fused), as in tests/test_dropin_live.py.
"""
import numpy as np
import pytest

import fused_host
import irl_control_b200 as pkg
import irl_control_b200.osc as pkg_osc
from conftest import GOLDEN_CASES, GOLDEN_CASES_F4, load_golden
from irl_control_b200 import configs
from irl_control_b200.dual_ur5 import DualUR5Model
from irl_control_b200.synthetic import SCENARIOS, patched_config
from oracle.ref_harness import FakeSim


class _Sim(FakeSim):
    def full_mass_matrix(self):                      # what `_mj_fullM` yields (robot.py:69-70)
        nv = self.model.nv
        return np.asarray(self.data.qM, dtype=np.float64).reshape(nv, nv)


class _HostEngine:
    def __init__(self, layout, device=None):
        self.layout, self.k, self.n_ctrl = layout, layout.k, layout.n_ctrl

    def step_host(self, state, **kw):
        st = dict(state)
        if "ft_xmat" in st:
            st["ft_xmat"] = st["ft_xmat"].reshape(st["ft_xmat"].shape[0], -1, 9)
        return fused_host.run_stream(self.layout, st)


@pytest.mark.parametrize("case", GOLDEN_CASES + GOLDEN_CASES_F4)
def test_generate_reproduces_the_reference_goldens(case, monkeypatch):
    monkeypatch.setattr(pkg_osc, "BatchedOSC", _HostEngine)
    generate_on_golden(case)


def generate_on_golden(case):
    """Shared with tests/test_gpu_zz_dropin.py, which runs it with the real engine (libirlosc.so on the GPU)."""
    g, ld = load_golden(case)
    sc = SCENARIOS[str(g["scenario"])]
    cfg = patched_config(sc)
    model = DualUR5Model(n_free_objects=configs.SCENE_FREE_OBJECTS[sc["scene"]])
    sim = _Sim(model)
    devices = [pkg.Device(d, model, sim, True) for d in cfg["devices"]]
    robot = pkg.Robot([devices[i] for i in cfg["robots"][0]["device_ids"]], "DualUR5", sim, True)
    by_name = {c["name"]: c for c in cfg["controller_configs"]}
    osc = pkg.OSC(robot, sim, [(dev, dict(by_name[c])) for dev, c in sc["device_cfgs"]],
                  dict(by_name["nullspace"]) if ld["nullspace_kv"] is not None else None,
                  use_g=ld["use_g"], admittance=sc["admittance"])
    names = list(sc["targets"])
    B = g["dq"].shape[0]
    keys = ("M", "J6", "dq", "bias", "ee_xyz", "ee_quat", "ft_xmat", "ft_raw")
    worst = 0.0
    for i in range(B):
        sim.load_instance({k: g[k][i] for k in keys}, names, devices)
        targets = {}
        for d, nm in enumerate(names):
            t = pkg.Target(np.zeros(6), np.zeros(6))
            t.set_xyz(g["target_xyz"][i][d])
            t.set_quat(g["target_quat"][i][d])
            tv = g["target_vel"][i][d]
            if np.any(tv != 0):                       # stored as osc.py:172 saw them: [xyz_vel, abg_vel]
                t.set_xyz_vel(tv[:3])
                t.set_abg_vel(tv[3:])
            targets[nm] = t
            robot.get_device(nm).max_vel = ([float(g["max_vel"][i][d][0]), float(g["max_vel"][i][d][1])]
                                            if ld["devices"][d]["has_max_vel"] else None)
        if g["index_error"][i]:
            with pytest.raises(IndexError):          # what the reference raised for this instance
                osc.generate(targets)
            continue
        idxs, forces = osc.generate(targets)
        got = np.concatenate(forces)
        worst = max(worst, np.abs(got - g["ctrl"][i]).max() / np.abs(g["u_all"][i]).max())
        assert [list(x) for x in idxs] == [list(robot.get_device(nm).ctrl_idxs) for nm in names]
    assert worst < 1e-6, worst


def test_generate_from_the_polling_thread_cache(monkeypatch):
    """`use_sim=False` (robot.py:103-123): `Robot.start()` runs in a thread, `generate` is served from the cached
    snapshot - M, J, dq, EE poses and, with admittance, the ROTATED wrench of that same snapshot (osc.py:179) - and
    returns what the live-simulator mode returns for the same state."""
    import threading
    import time
    monkeypatch.setattr(pkg_osc, "BatchedOSC", _HostEngine)
    g, ld = load_golden("admit_test_s1")
    sc = SCENARIOS["admit_test"]
    cfg = patched_config(sc)
    model = DualUR5Model(n_free_objects=configs.SCENE_FREE_OBJECTS[sc["scene"]])
    names = list(sc["targets"])
    by_name = {c["name"]: c for c in cfg["controller_configs"]}
    keys = ("M", "J6", "dq", "bias", "ee_xyz", "ee_quat", "ft_xmat", "ft_raw")
    results = {}
    for use_sim in (True, False):
        sim = _Sim(model)
        devices = [pkg.Device(d, model, sim, use_sim) for d in cfg["devices"]]
        robot = pkg.Robot([devices[i] for i in cfg["robots"][0]["device_ids"]], "DualUR5", sim, use_sim)
        osc = pkg.OSC(robot, sim, [(dev, dict(by_name[c])) for dev, c in sc["device_cfgs"]], dict(by_name["nullspace"]),
                      admittance=True)
        sim.load_instance({k: g[k][3] for k in keys}, names, devices)
        targets = {}
        for d, nm in enumerate(names):
            t = pkg.Target(np.zeros(6), np.zeros(6))
            t.set_xyz(g["target_xyz"][3][d])
            t.set_quat(g["target_quat"][3][d])
            targets[nm] = t
        th = None
        if not use_sim:
            with pytest.raises(AssertionError):
                osc.generate(targets)                           # not running yet (osc.py:129-130)
            th = threading.Thread(target=robot.start, daemon=True)
            th.start()
            deadline = time.time() + 5.0
            while time.time() < deadline and len(robot._cache) < 3:
                time.sleep(0.005)
            time.sleep(0.02)                                    # at least one full refresh of every variable
            assert robot.is_running()
        results[use_sim] = np.concatenate(osc.generate(targets)[1])
        if th is not None:
            robot.stop()
            th.join(timeout=2.0)
            assert not th.is_alive()
    assert np.abs(results[True] - g["ctrl"][3]).max() < 1e-6 * np.abs(g["u_all"][3]).max()
    assert np.allclose(results[False], results[True], rtol=0, atol=1e-9 * np.abs(results[True]).max())
