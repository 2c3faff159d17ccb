"""TEST INFRASTRUCTURE ONLY - run the UNMODIFIED reference OSC against a fake sim.

Imports `irl_control/{device,robot,osc,utils}.py` straight from
/root/reference (or its verbatim, git-ignored copy oracle/_ref/ staged by oracle/stage_ref.py) after registering stub modules for
the two third-party packages those files import and that do not exist in this
image:

    mujoco_py      (osc.py:3, robot.py:3)  -> only `cymj._mj_fullM` is called
                                              (robot.py:69); `load_model_from_path`
                                              / `MjSim` are import-time names
                                              for mujoco_app.py:2
    transforms3d   (osc.py:4-7, utils.py:3) -> oracle/t3d.py

`FakeSim` serves one robot instance's state from arrays, exposing exactly the
`sim.model` / `sim.data` accessors the reference touches.
This module only works where /root/reference exists (this container); it is
used by tests/golden/make_golden.py and by tests that are skipped elsewhere.
"""
from __future__ import annotations

import os
import sys
import types
from typing import Dict, List, Optional, Sequence

import numpy as np
import yaml

def _reference_root() -> str:
    """/root/reference (build container), else the verbatim copy staged by oracle/stage_ref.py (GPU box)."""
    env = os.environ.get("IRL_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile("/root/reference/irl_control/osc.py"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REFERENCE_ROOT = _reference_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "irl_control", "osc.py"))


def _install_stubs():
    from . import t3d
    if "mujoco_py" not in sys.modules:
        mjp = types.ModuleType("mujoco_py")
        cymj = types.ModuleType("mujoco_py.cymj")

        def _mj_fullM(model, dst, qM):
            dst[:] = np.asarray(qM).reshape(-1)

        cymj._mj_fullM = _mj_fullM
        mjp.cymj = cymj
        mjp.load_model_from_path = lambda *a, **k: (_ for _ in ()).throw(
            RuntimeError("mujoco_py stub: no simulator in this image"))
        mjp.MjSim = object
        sys.modules["mujoco_py"] = mjp
        sys.modules["mujoco_py.cymj"] = cymj
    if "transforms3d" not in sys.modules:
        pkg = types.ModuleType("transforms3d")
        der = types.ModuleType("transforms3d.derivations")
        derq = types.ModuleType("transforms3d.derivations.quaternions")
        derq.qmult = t3d.qmult
        qs = types.ModuleType("transforms3d.quaternions")
        qs.qconjugate = t3d.qconjugate
        qs.qmult = lambda a, b: np.array(t3d.qmult(a, b))
        qs.quat2mat = t3d.quat2mat
        eu = types.ModuleType("transforms3d.euler")
        eu.quat2euler = t3d.quat2euler
        eu.euler2quat = t3d.euler2quat
        eu.mat2euler = t3d.mat2euler
        eu.euler2mat = t3d.euler2mat
        eu.quat2mat = t3d.quat2mat
        ut = types.ModuleType("transforms3d.utils")
        ut.normalized_vector = t3d.normalized_vector
        pkg.derivations, pkg.quaternions, pkg.euler, pkg.utils = der, qs, eu, ut
        der.quaternions = derq
        for name, mod in [("transforms3d", pkg), ("transforms3d.derivations", der),
                          ("transforms3d.derivations.quaternions", derq),
                          ("transforms3d.quaternions", qs), ("transforms3d.euler", eu),
                          ("transforms3d.utils", ut)]:
            sys.modules[name] = mod


def import_reference():
    """Returns the reference's (Device, Robot, OSC, Target, DeviceState, RobotState) classes."""
    if not reference_available():
        raise RuntimeError("reference sources not found under %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import irl_control  # noqa: F401  (pulls device, robot, osc, mujoco_app)
    from irl_control.device import Device, DeviceState
    from irl_control.robot import Robot, RobotState
    from irl_control.osc import OSC
    from irl_control.utils import Target
    return Device, Robot, OSC, Target, DeviceState, RobotState


def load_reference_yaml(name: str, inject_start_body: bool = True) -> Dict:
    """Reads robot_configs/<name> from the reference.  SURVEY.md N1: the default
    YAMLs comment out `start_body`, which makes `Device.__init__` fail for the
    arms (7 joints vs 6 start angles); `inject_start_body` restores the
    `iros2022.yaml:13,22` setting so the DualUR5 can be constructed."""
    path = os.path.join(REFERENCE_ROOT, "irl_control", "robot_configs", name)
    with open(path, "r") as fh:
        cfg = yaml.safe_load(fh)
    if inject_start_body:
        for dev in cfg["devices"]:
            if dev["name"] in ("ur5right", "ur5left"):
                dev.setdefault("start_body", "dual_ur_stand")
    return cfg


class _FakeData:
    def __init__(self, model):
        self._model = model
        self.qpos = np.zeros(model.nq)
        self.qvel = np.zeros(model.nv)
        self.qacc = np.zeros(model.nv)
        self.qM = np.eye(model.nv).reshape(-1)
        self.qfrc_bias = np.zeros(model.nv)
        self.sensordata = np.zeros(model.nsensordata)
        self.ctrl = np.zeros(model.nu)
        self.xpos: Dict[str, np.ndarray] = {}
        self.xquat: Dict[str, np.ndarray] = {}
        self.xvelp: Dict[str, np.ndarray] = {}
        self.jacp: Dict[str, np.ndarray] = {}
        self.jacr: Dict[str, np.ndarray] = {}
        self.site_xmat: Dict[str, np.ndarray] = {}

    def get_body_xpos(self, name):
        return self.xpos[name]

    def get_body_xquat(self, name):
        return self.xquat[name]

    def get_body_xvelp(self, name):
        return self.xvelp.get(name, np.zeros(3))

    def get_body_jacp(self, name):
        return self.jacp[name].reshape(-1)

    def get_body_jacr(self, name):
        return self.jacr[name].reshape(-1)

    def get_site_xmat(self, name):
        return self.site_xmat[name]


class FakeSim:
    """`sim` with `.model` and `.data`; state is loaded from arrays."""

    def __init__(self, model):
        self.model = model
        self.data = _FakeData(model)

    def forward(self):
        pass

    def load_instance(self, st: Dict[str, np.ndarray], target_names: Sequence[str], devices):
        """st holds ONE instance with per-device arrays in TARGET order:
        M (n,n), J6 (D,6,n), dq (n), bias (n), ee_xyz (D,3), ee_quat (D,4),
        ft_xmat (D,9), ft_raw (D,6).  Sub-devices that are not targeted still get
        pulled by `Robot.get_all_states` (robot.py:130-131); they are served zeros,
        which cannot reach the output (their Jacobian is never stacked, osc.py:137-138)."""
        m, d = self.model, self.data
        n, nv = m.nv_robot, m.nv
        Mfull = np.eye(nv) * 0.05
        Mfull[:n, :n] = st["M"]
        d.qM = Mfull.reshape(-1)
        d.qvel[:] = 0.0
        d.qvel[:n] = st["dq"]
        d.qfrc_bias[:] = 0.0
        d.qfrc_bias[:n] = st["bias"]
        d.sensordata[:] = 0.0
        slices = {"ur5right": 0, "ur5left": 6}      # dual_ur5.xml:289-293 via device.py:150-167
        for dev in devices:
            body = dev.EE
            jp = np.zeros((3, nv))
            jr = np.zeros((3, nv))
            d.xpos[body] = np.zeros(3)
            d.xquat[body] = np.array([1.0, 0.0, 0.0, 0.0])
            if dev.name in slices:
                d.site_xmat["ft_frame_" + dev.name] = np.eye(3)
            if dev.name in target_names:
                i = list(target_names).index(dev.name)
                jp[:, :n] = st["J6"][i, :3]
                jr[:, :n] = st["J6"][i, 3:]
                d.xpos[body] = np.asarray(st["ee_xyz"][i], dtype=np.float64)
                d.xquat[body] = np.asarray(st["ee_quat"][i], dtype=np.float64)
                if dev.name in slices:
                    d.site_xmat["ft_frame_" + dev.name] = np.asarray(st["ft_xmat"][i]).reshape(3, 3)
                    o = slices[dev.name]
                    d.sensordata[o:o + 6] = st["ft_raw"][i]
            d.jacp[body], d.jacr[body] = jp, jr


class ReferenceRunner:
    """Builds the reference's Device/Robot/OSC once and evaluates `OSC.generate`
    instance by instance on supplied state arrays."""

    def __init__(self, model, robot_cfg: Dict, osc_device_cfgs: Sequence, target_names: Sequence[str],
                 nullspace_cfg: Optional[Dict], use_g: bool = True, admittance: bool = False):
        Device, Robot, OSC, Target, DeviceState, RobotState = import_reference()
        self.Target = Target
        self.sim = FakeSim(model)
        self.devices = [Device(d, model, self.sim, True) for d in robot_cfg["devices"]]
        ids = robot_cfg["robots"][0]["device_ids"]
        self.robot = Robot([self.devices[i] for i in ids], robot_cfg["robots"][0]["name"], self.sim, True)
        cfg_by_name = {c["name"]: c for c in robot_cfg["controller_configs"]}
        # the reference mutates these dicts (osc.py:38-39): give it private copies
        dev_cfgs = [(name, dict(cfg_by_name[cfg_name])) for name, cfg_name in osc_device_cfgs]
        ns = dict(cfg_by_name[nullspace_cfg]) if nullspace_cfg else None
        self.osc = OSC(self.robot, self.sim, dev_cfgs, ns, use_g=use_g, admittance=admittance)
        self.target_names = list(target_names)
        self.n = self.robot.num_joints_total

    def run(self, st: Dict[str, np.ndarray], tgt_xyz, tgt_quat, tgt_vel=None, max_vel=None):
        """tgt_*: arrays indexed like `target_names`.  Returns dict with the packed forces
        (list per target), ctrl index lists, the full joint vector u_all and the branch flag."""
        self.sim.load_instance(st, self.target_names, self.devices)
        targets = {}
        for i, name in enumerate(self.target_names):
            t = self.Target()
            t.set_xyz(np.array(tgt_xyz[i], dtype=np.float64))
            t.set_quat(np.array(tgt_quat[i], dtype=np.float64))
            if tgt_vel is not None:
                t.set_xyz_vel(np.array(tgt_vel[i][:3], dtype=np.float64))
                t.set_abg_vel(np.array(tgt_vel[i][3:], dtype=np.float64))
            targets[name] = t
            if max_vel is not None:
                # NaN in the first slot stands for `device.max_vel = None` (osc.py:163: the un-limited gain branch)
                self.robot.get_device(name).max_vel = (None if np.isnan(max_vel[i][0]) else
                                                       [float(max_vel[i][0]), float(max_vel[i][1])])
        calls = {"pinv": 0}
        real_pinv = np.linalg.pinv

        def counting_pinv(*a, **k):
            calls["pinv"] += 1
            return real_pinv(*a, **k)

        np.linalg.pinv = counting_pinv
        try:
            idxs, forces = self.osc.generate(targets)
            # second pass with every joint "actuated" to expose the internal u_all
            saved = {}
            first = self.target_names[0]
            dev0 = self.robot.get_device(first)
            saved = dev0.actuator_trnids
            dev0.actuator_trnids = np.arange(self.n)
            try:
                _, full = self.osc.generate(targets)
            finally:
                dev0.actuator_trnids = saved
        finally:
            np.linalg.pinv = real_pinv
        return {
            "ctrl_idxs": [np.asarray(i) for i in idxs],
            "forces": [np.asarray(f, dtype=np.float64) for f in forces],
            "u_all": np.asarray(full[0], dtype=np.float64),
            "pinv": calls["pinv"] > 0,
            # what osc.py:172 actually sees: abg_vel goes Euler -> quaternion -> Euler inside Target
            "target_vel_seen": np.stack([np.hstack([targets[nm].get_xyz_vel(), targets[nm].get_abg_vel()])
                                         for nm in self.target_names]),
        }

    def calc_error(self, st, name, tgt_xyz, tgt_quat):
        self.sim.load_instance(st, self.target_names, self.devices)
        t = self.Target()
        t.set_xyz(np.array(tgt_xyz, dtype=np.float64))
        t.set_quat(np.array(tgt_quat, dtype=np.float64))
        return self.osc.calc_error(t, self.robot.get_device(name))


def scenario_runner(scenario: str, use_g: bool = True, nullspace: bool = True) -> "ReferenceRunner":
    """The reference's own Device / Robot / OSC for a named scenario of irl_control_b200.synthetic.SCENARIOS, built from
    the reference's YAML (start_body injected) on the fake simulator."""
    from irl_control_b200.configs import SCENE_FREE_OBJECTS
    from irl_control_b200.dual_ur5 import DualUR5Model
    from irl_control_b200.synthetic import SCENARIOS
    sc = SCENARIOS[scenario]
    cfg = load_reference_yaml(sc["config"].replace("+start_body", ""), inject_start_body=True)
    for dev in cfg["devices"]:                          # a scenario may stand for a user-edited YAML
        dev.update(sc.get("config_patch", {}).get(dev["name"], {}))
    model = DualUR5Model(n_free_objects=SCENE_FREE_OBJECTS[sc["scene"]])
    return ReferenceRunner(model, cfg, sc["device_cfgs"], sc["targets"], "nullspace" if nullspace else None,
                           use_g=use_g, admittance=sc["admittance"])


def time_reference_generate(runner: "ReferenceRunner", batch: Dict[str, np.ndarray], repeat: int = 1):
    """Seconds spent inside the UNMODIFIED `OSC.generate` (osc.py:120-210, state pulls of robot.py / device.py
    included) over the instances of `batch` (oracle field names, see synthetic.oracle_inputs); loading an instance
    into the fake simulator is not timed.  Returns (seconds, calls, last forces)."""
    import time
    B = batch["M"].shape[0]
    spent, calls, forces = 0.0, 0, None
    for _ in range(repeat):
        for i in range(B):
            st = {"M": batch["M"][i], "J6": batch["J"][i], "dq": batch["dq"][i], "bias": batch["bias"][i],
                  "ee_xyz": batch["ee_xyz"][i], "ee_quat": batch["ee_quat"][i], "ft_xmat": batch["ft_xmat"][i],
                  "ft_raw": batch["ft_raw"][i]}
            runner.sim.load_instance(st, runner.target_names, runner.devices)
            targets = {}
            for d, name in enumerate(runner.target_names):
                t = runner.Target()
                t.set_xyz(np.array(batch["tgt_xyz"][i][d], dtype=np.float64))
                t.set_quat(np.array(batch["tgt_quat"][i][d], dtype=np.float64))
                targets[name] = t
                runner.robot.get_device(name).max_vel = [float(batch["max_vel"][i][d][0]), float(batch["max_vel"][i][d][1])]
            t0 = time.perf_counter()
            _, forces = runner.osc.generate(targets)
            spent += time.perf_counter() - t0
            calls += 1
    return spent, calls, forces


# ---------------------------------------------------------------- the insertion demo's caller loop, unmodified
def import_insertion_task():
    """`irl_control.examples.insertion_task` (unmodified) with stubs for what it imports beyond the OSC path:
    `mujoco_py.mjviewer.MjViewer` (never constructed here) and `transforms3d.affines.compose`."""
    import_reference()
    from . import t3d
    mjp = sys.modules["mujoco_py"]
    if "mujoco_py.mjviewer" not in sys.modules:
        mv = types.ModuleType("mujoco_py.mjviewer")
        mv.MjViewer = object
        mjp.mjviewer = mv
        sys.modules["mujoco_py.mjviewer"] = mv
    if "transforms3d.affines" not in sys.modules:
        af = types.ModuleType("transforms3d.affines")

        def compose(T, R, Z, S=None):                 # transforms3d.affines.compose without shear
            A = np.eye(4)
            A[:3, :3] = np.asarray(R) @ np.diag(np.asarray(Z, dtype=np.float64))
            A[:3, 3] = T
            return A

        af.compose = compose
        sys.modules["transforms3d"].affines = af
        sys.modules["transforms3d.affines"] = af
    # insertion_task.py:19 evaluates `quat2euler(..., 'rxyz')` once at import for a constant nothing reads
    # (DEFAULT_EE_ORIENTATION); the restatement covers the default 'sxyz' axes only, so that single call gets NaNs
    eu = sys.modules["transforms3d.euler"]
    plain = eu.quat2euler
    eu.quat2euler = lambda q, axes='sxyz': plain(q) if axes == 'sxyz' else (float("nan"),) * 3
    try:
        import irl_control.examples.insertion_task as mod
    finally:
        eu.quat2euler = plain
    return mod


class SequenceStop(Exception):
    pass


def drive_reference_sequence(actions, action_objects, object_qpos, poses, active, ctrlr_dof, max_vel0, n_ticks,
                             step_period):
    """Runs the reference's own `run_sequence` / `go_to_waypoint` / `grip` / `send_forces` /
    `set_waypoint_targets` (insertion_task.py, unmodified methods) on a pose stream instead of a simulator.

    object_qpos[joint_name] = 7-vector (what `sim.data.get_joint_qpos` returns); poses: active_xyz / active_quat /
    passive_xyz indexed by tick; GRIP durations are turned into step budgets of round(duration / step_period)
    (the reference waits on a wall-clock timer thread).  Returns one record per `controller.generate` call."""
    mod = import_insertion_task()
    Device, Robot, OSC, Target, DeviceState, RobotState = import_reference()
    passive = "ur5left" if active == "ur5right" else "ur5right"
    tick = [0]
    rec = []

    class FakeDevice:
        def __init__(self, name, is_active):
            self.name, self._active = name, is_active
            self.max_vel = [float(max_vel0), 5.0]
            self.ctrlr_dof_xyz, self.ctrlr_dof_abg = list(ctrlr_dof[:3]), list(ctrlr_dof[3:])

        def get_state(self, which):
            t = tick[0]
            if which == DeviceState.EE_XYZ:
                return np.array(poses["active_xyz"][t] if self._active else poses["passive_xyz"][t], dtype=np.float64)
            if which == DeviceState.EE_QUAT:
                assert self._active
                return np.array(poses["active_quat"][t], dtype=np.float64)
            raise KeyError(which)

    devs = {active: FakeDevice(active, True), passive: FakeDevice(passive, False)}

    class FakeData:
        ctrl = np.zeros(15)

        @staticmethod
        def get_joint_qpos(name):
            return np.array(object_qpos[name], dtype=np.float64)

    class FakeSim:
        data = FakeData()

        @staticmethod
        def step():
            tick[0] += 1
            if tick[0] >= n_ticks:
                raise SequenceStop

    class FakeController:
        @staticmethod
        def generate(targets):
            rec.append(dict(tick=tick[0], action=task._cur, err=float(task.errors.get(active, 0.0)),
                            max_vel0=float(devs[active].max_vel[0]),
                            active_xyz=np.array(targets[active].get_xyz(), dtype=np.float64),
                            active_quat=np.array(targets[active].get_quat(), dtype=np.float64),
                            passive_xyz=np.array(targets[passive].get_xyz(), dtype=np.float64),
                            passive_quat=np.array(targets[passive].get_quat(), dtype=np.float64)))
            return [np.arange(1, 8)], [np.zeros(7)]

        @staticmethod
        def calc_error(target, device):
            return OSC.calc_error(None, target, device)           # the reference's own method (uses no self state)

    class Driven(mod.InsertionTask):
        def __init__(self):                                         # the real one needs a simulator and a viewer
            pass

        def sleep_for(self, sleep_time):                           # the timer thread's body: replaced by a step budget
            pass

        @property
        def timer_running(self):
            if self._budget > 0:
                self._budget -= 1
                return True
            return False

        def go_to_waypoint(self, params):
            self._cur += 1
            return super().go_to_waypoint(params)

        def grip(self, params):
            self._cur += 1
            self._budget = int(round(float(params['gripper_duration']) / step_period))
            return super().grip(params)

        def send_forces(self, forces, gripper_force=None, update_errors=None, render=True):
            rec[-1]["gripper_force"] = float(gripper_force) if gripper_force else 0.0
            return super().send_forces(forces, gripper_force=gripper_force, update_errors=update_errors, render=render)

    task = Driven()
    task._cur, task._budget = -1, 0
    task.ur5right, task.ur5left = devs["ur5right"], devs["ur5left"]
    task.set_active_arm("right" if active == "ur5right" else "left")
    task.sim = FakeSim()
    task.viewer = types.SimpleNamespace(render=lambda: None)
    task.controller = FakeController()
    task.robot = types.SimpleNamespace(get_device=lambda name: devs[name])
    task.errors = {}
    task.action_objects = action_objects
    task.action_map = task.get_action_map()
    task.DEFAULT_PARAMS = dict((a, task.get_default_action_ctrl_params(a)) for a in mod.Action)
    task.targets = {active: Target(), passive: Target()}
    try:
        task.run_sequence(actions)
    except SequenceStop:
        pass
    return rec


def reference_object_placement(action_objects, arm_name=None, seed=None):
    """The reference's `initialize_action_objects` (arm_name None) or `initialize_action_objects_random`
    (insertion_task.py:299-311, 341-369), unmodified, with `set_free_joint_qpos` captured instead of written into a
    simulator.  Returns {joint_name: (pos, quat)} and, for the random variant, the six numbers `np.random.uniform`
    produced (same seed replayed in the reference's call order)."""
    mod = import_insertion_task()
    placed = {}

    class Driven(mod.InsertionTask):
        def __init__(self):
            pass

        def set_free_joint_qpos(self, free_joint_name, quat=None, pos=None):
            placed[free_joint_name] = (np.array(pos, dtype=np.float64), np.array(quat, dtype=np.float64))

    task = Driven()
    task.action_objects = action_objects
    if arm_name is None:
        task.initialize_action_objects()
        return placed, None
    np.random.seed(seed)
    task.initialize_action_objects_random(arm_name)
    np.random.seed(seed)
    draws = [np.random.uniform(lo, hi) for lo, hi in ((0.4, 0.6), (0.5, 0.7), (0.0, 0.3), (0.5, 0.7), (-20, 20), (-20, 20))]
    return placed, draws


def drive_reference_gain_test(right_wps, left_wps, ee_right, ee_left, n_ticks):
    """The reference's `GainTest.run('gain_test', ...)` (examples/gain_test.py:98-172, unmodified) on a pose stream:
    `ee_right[t]` / `ee_left[t]` are the EE positions the t-th loop iteration sees.  The demo timer becomes a
    budget of n_ticks iterations.  Returns (right target, left target, right index, left index) per `generate`."""
    import_insertion_task()                                      # registers the mjviewer stub
    Device, Robot, OSC, Target, DeviceState, RobotState = import_reference()
    mjp = sys.modules["mujoco_py"]
    if not hasattr(mjp, "GlfwContext"):
        mjp.GlfwContext = object
    import irl_control.examples.gain_test as mod
    tick = [0]
    rec = []
    right_wps, left_wps = np.asarray(right_wps, dtype=np.float64), np.asarray(left_wps, dtype=np.float64)

    class FakeDevice:
        def __init__(self, stream):
            self._stream = stream

        def get_state(self, which):
            assert which == DeviceState.EE_XYZ
            return np.array(self._stream[tick[0]], dtype=np.float64)

    devs = {"ur5right": FakeDevice(ee_right), "ur5left": FakeDevice(ee_left)}

    def index_of(wps, xyz):
        return int(np.argmin(np.abs(wps - np.asarray(xyz)).sum(axis=1)))

    class FakeController:
        @staticmethod
        def generate(targets):
            r, l = np.array(targets["ur5right"].get_xyz()), np.array(targets["ur5left"].get_xyz())
            rec.append((r, l, index_of(right_wps, r), index_of(left_wps, l)))
            return [np.arange(1, 8)], [np.zeros(7)]

    class FakeData:
        ctrl = np.zeros(15)

        @staticmethod
        def set_mocap_pos(name, pos):
            pass

    class Driven(mod.GainTest):
        def __init__(self):
            self._budget = n_ticks

        def sleep_for(self, sleep_time):
            pass

        @property
        def timer_running(self):
            if self._budget > 0:
                self._budget -= 1
                return True
            return False

        def gen_waypoint_path(self):
            return right_wps, left_wps

    task = Driven()
    task.robot = types.SimpleNamespace(get_device=lambda name: devs[name], stop=lambda: None)
    task.robot_data_thread = types.SimpleNamespace(join=lambda: None)
    task.controller = FakeController()
    task.errors = {}
    task.viewer = types.SimpleNamespace(render=lambda: None)
    task.sim = types.SimpleNamespace(data=FakeData(), step=lambda: tick.__setitem__(0, tick[0] + 1))
    task.run("gain_test", 10)
    return rec
