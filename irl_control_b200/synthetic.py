"""Seeded synthetic DualUR5 workloads for the parity tests and the benchmark.

A scenario fixes what the reference's demo constructors fix (robot YAML,
controller configs per device, target order, admittance flag - SURVEY.md
section 8 table) and `synth_batch` draws B physically consistent states
through the rigid-body model in `dual_ur5.py`:

    q  ~ U(-pi, pi) on stand + arm joints, U(0, 0.8) on gripper joints
    dq ~ N(0, 0.3^2)
    M, J, qfrc_bias, EE pose, F/T frame  <- dynamics(q, dq)
    target pose = EE pose perturbed by N(0, 0.1^2) m and U(-0.5, 0.5) rad (Euler)
    F/T sensor  ~ N(0, 5^2) N, N(0, 0.5^2) N m              (admittance scenarios)
    max_vel[0]  = clip(6 * |err|, 0.1, 3.0) on the active arm (insertion scenario,
                  insertion_task.py:91-95,294-295)

This is input synthesis; nothing here is timed or is part of the control law.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .configs import IROS2022_DEVICE_CONFIG, device_cfgs, robot_config
from .dual_ur5 import DualUR5Model, dynamics, sample_joint_states
from .layout import OscLayout

SCENARIOS: Dict[str, Dict] = {
    # gain_test.py:28-36,124-128,180
    "gain_test": dict(config="default_xyz.yaml+start_body", scene="gain_test_scene.xml",
                      device_cfgs=[("base", "osc0"), ("ur5right", "osc2"), ("ur5left", "osc2")],
                      targets=["ur5right", "ur5left", "base"], admittance=False),
    # admit_test.py:19-25,55-58,85
    "admit_test": dict(config="default_xyz_abg.yaml+start_body", scene="admit_test_scene.xml",
                       device_cfgs=[("ur5right", "osc2"), ("ur5left", "osc2")],
                       targets=["ur5right", "ur5left"], admittance=True),
    # insertion_task.py:52-60,77-80,424 (base configured but not targeted)
    "insertion": dict(config="default_xyz_abg.yaml+start_body", scene="insertion_task_scene.xml",
                      device_cfgs=[("base", "osc0"), ("ur5right", "osc2"), ("ur5left", "osc2")],
                      targets=["ur5right", "ur5left"], admittance=False),
    # space_mouse_example.py:106-110 shape: three 6-DoF-capable devices, k = 13
    "worst_case": dict(config="default_xyz_abg.yaml+start_body", scene="gain_test_scene.xml",
                       device_cfgs=[("base", "osc0"), ("ur5right", "osc2"), ("ur5left", "osc2")],
                       targets=["ur5right", "ur5left", "base"], admittance=False),
    # admittance=True with the base among the targets (no example does): the base has no F/T sensor, its wrench is
    # identically zero (device.py:150-170), the arms' wrenches enter as in admit_test
    "worst_case_admit": dict(config="default_xyz_abg.yaml+start_body", scene="gain_test_scene.xml",
                             device_cfgs=[("base", "osc0"), ("ur5right", "osc2"), ("ur5left", "osc2")],
                             targets=["ur5right", "ur5left", "base"], admittance=True),
    # a user-edited YAML: DoF masks no shipped config has (right arm xyz + b, g; left arm xyz + a) -> 5 + 4 + 1 rows.
    # No specialised kernel serves unequal arm row counts: this is the generic kernel's case.
    "mixed_dof": dict(config="default_xyz_abg.yaml+start_body", scene="gain_test_scene.xml",
                      device_cfgs=[("base", "osc0"), ("ur5right", "osc2"), ("ur5left", "osc2")],
                      targets=["ur5right", "ur5left", "base"], admittance=False,
                      config_patch={"ur5right": {"ctrlr_dof_abg": [False, True, True]},
                                    "ur5left": {"ctrlr_dof_abg": [True, False, False]}}),
    # robot_configs/iros2022.yaml (osc0 = osc2 = kp 200 / kv 20 / ko 75, base max_vel [0, 2],
    # arms [2, 5], start_body set) with the devices, controllers and order of iros2022_task.yaml:1-4
    "iros2022": dict(config="iros2022.yaml", scene="iros2022.xml",
                     device_cfgs=device_cfgs(IROS2022_DEVICE_CONFIG),
                     targets=list(IROS2022_DEVICE_CONFIG["devices"]), admittance=False),
}


def patched_config(sc: Dict) -> Dict:
    """The scenario's robot config with its optional per-device `config_patch` applied (a user-edited YAML)."""
    cfg = robot_config(sc["config"])
    for dev in cfg["devices"]:
        dev.update(sc.get("config_patch", {}).get(dev["name"], {}))
    return cfg


def build_scenario(name: str):
    """(app, osc, target_names, layout) for a named scenario, built through the host API."""
    from .mujoco_app import MujocoApp
    from .osc import OSC
    sc = SCENARIOS[name]
    app = MujocoApp(sc["config"], sc["scene"], config_override=patched_config(sc) if "config_patch" in sc else None)
    robot = app.get_robot("DualUR5")
    cfgs = [(dev, app.get_controller_config(cfg)) for dev, cfg in sc["device_cfgs"]]
    osc = OSC(robot, app.sim, cfgs, app.get_controller_config("nullspace"), admittance=sc["admittance"])
    return app, osc, list(sc["targets"]), osc.layout_for(sc["targets"])


def scenario_layout(name: str) -> OscLayout:
    return build_scenario(name)[3]


# ---------------------------------------------------------------- batched quaternion helpers
def _euler_to_quat(e: torch.Tensor) -> torch.Tensor:
    h = 0.5 * e
    ci, cj, ck = torch.cos(h[..., 0]), torch.cos(h[..., 1]), torch.cos(h[..., 2])
    si, sj, sk = torch.sin(h[..., 0]), torch.sin(h[..., 1]), torch.sin(h[..., 2])
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    return torch.stack([cj * cc + sj * ss, cj * sc - sj * cs, cj * ss + sj * cc, cj * cs - sj * sc], -1)


def _quat_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack([aw * bw - ax * bx - ay * by - az * bz,
                        aw * bx + ax * bw + ay * bz - az * by,
                        aw * by + ay * bw + az * bx - ax * bz,
                        aw * bz + az * bw + ax * by - ay * bx], -1)


def pack_lower(M: torch.Tensor) -> torch.Tensor:
    """(B, n, n) symmetric -> (B, n(n+1)/2) row-major lower triangle (IRLOSC_M_PACKED)."""
    n = M.shape[-1]
    i, j = torch.tril_indices(n, n, device=M.device)
    return M[:, i, j].contiguous()


def synth_batch(layout: OscLayout, B: int, seed: int = 0, device="cpu", model: Optional[DualUR5Model] = None,
                per_instance_max_vel: bool = True, insertion_schedule: bool = False,
                chunk: int = 16384) -> Dict[str, torch.Tensor]:
    """Draw B instances for `layout` (per-device fields in target order).  All float64 on `device`."""
    model = model or DualUR5Model()
    dev = torch.device(device)
    names = [d.name for d in layout.devices]
    ee_body = {"base": "ur_stand_dummy", "ur5right": "ur_EE_ur5right", "ur5left": "ur_EE_ur5left"}
    ft_site = {"ur5right": "ft_frame_ur5right", "ur5left": "ft_frame_ur5left"}
    q_np, dq_np = sample_joint_states(B, seed)
    rng = np.random.default_rng(seed + 7919)
    D = len(names)
    dxyz = rng.normal(0.0, 0.1, size=(B, D, 3))
    deul = rng.uniform(-0.5, 0.5, size=(B, D, 3))
    ft_raw = np.concatenate([rng.normal(0.0, 5.0, size=(B, D, 3)), rng.normal(0.0, 0.5, size=(B, D, 3))], -1)
    out: Dict[str, List[torch.Tensor]] = {k: [] for k in
                                          ("M", "J6", "q", "dq", "bias", "ee_xyz", "ee_quat", "ft_xmat")}
    for s in range(0, B, chunk):
        q = torch.from_numpy(q_np[s:s + chunk]).to(dev)
        dq = torch.from_numpy(dq_np[s:s + chunk]).to(dev)
        dyn = dynamics(model, q, dq)
        j6, xyz, quat, fx = [], [], [], []
        for nm in names:
            b = model.body_name2id(ee_body[nm])
            jp, jr = dyn.jac_body(b)
            j6.append(torch.cat([jp, jr], dim=1))
            xyz.append(dyn.xpos[:, b])
            quat.append(dyn.xquat[:, b])
            if nm in ft_site:
                fx.append(dyn.site_xmat[:, model.site_name2id(ft_site[nm])].reshape(-1, 9))
            else:
                fx.append(torch.eye(3, dtype=torch.float64, device=dev).reshape(1, 9).expand(q.shape[0], 9))
        out["M"].append(dyn.M)
        out["q"].append(q)
        out["J6"].append(torch.stack(j6, 1))
        out["dq"].append(dq)
        out["bias"].append(dyn.bias)
        out["ee_xyz"].append(torch.stack(xyz, 1))
        out["ee_quat"].append(torch.stack(quat, 1))
        out["ft_xmat"].append(torch.stack(fx, 1))
    st = {k: torch.cat(v, 0).contiguous() for k, v in out.items()}
    rows = [(d, c) for d, dl in enumerate(layout.devices) for c in range(6) if dl.ctrlr_dof[c]]
    st["J"] = torch.stack([st["J6"][:, d, c] for d, c in rows], 1).contiguous()
    dxyz_t = torch.from_numpy(dxyz).to(dev)
    deul_t = torch.from_numpy(deul).to(dev)
    st["target_xyz"] = (st["ee_xyz"] + dxyz_t).contiguous()
    st["target_quat"] = _quat_mul(_euler_to_quat(deul_t), st["ee_quat"]).contiguous()
    st["ft_raw"] = torch.from_numpy(ft_raw).to(dev).contiguous()
    for nm_i, nm in enumerate(names):
        if nm not in ft_site:
            st["ft_raw"][:, nm_i] = 0.0          # device.py:162-163,169-170: no sensor -> zeros
    if per_instance_max_vel:
        mv = torch.tensor([list(d.max_vel) for d in layout.devices], dtype=torch.float64, device=dev)
        mv = mv[None].expand(B, D, 2).clone()
        if insertion_schedule:
            err = torch.cat([dxyz_t[:, 0], deul_t[:, 0]], -1).norm(dim=-1)
            mv[:, 0, 0] = torch.clamp(6.0 * err, 0.1, 3.0)
        st["max_vel"] = mv.contiguous()
    return st


def fused_inputs(st: Dict[str, torch.Tensor], layout: OscLayout, with_vel: bool = False) -> Dict[str, torch.Tensor]:
    """Fields `BatchedOSC.step_fused` consumes: joint states and targets only (no M / J / bias / EE)."""
    keep = {"q": st["q"], "dq": st["dq"], "target_xyz": st["target_xyz"], "target_quat": st["target_quat"]}
    if "max_vel" in st:
        keep["max_vel"] = st["max_vel"]
    if layout.admittance:
        keep["ft_raw"] = st["ft_raw"]
    if with_vel and "target_vel" in st:
        keep["target_vel"] = st["target_vel"]
    return keep


def scenario_model(name: str):
    """(layout, irlosc_model) of a named scenario for the fused step."""
    from .rigid_model import model_for_layout
    app, _osc, _names, layout = build_scenario(name)
    robot = app.get_robot("DualUR5")
    return layout, model_for_layout(app.sim.model, robot.joint_ids_all, layout)


def sparse_qM(M: torch.Tensor, joint_parent: Sequence[int], pad: int = 0) -> torch.Tensor:
    """(B, n, n) -> (B, nM + pad): the entries MuJoCo keeps in `mjData.qM` (IRLOSC_M_QM); `pad` extra doubles
    per instance stand for the free bodies' entries that follow the robot's in a scene."""
    from .layout import qm_index
    rows, cols = qm_index(joint_parent)
    q = M[:, torch.as_tensor(rows, device=M.device), torch.as_tensor(cols, device=M.device)]
    if pad:
        q = torch.cat([q, torch.full((M.shape[0], pad), float("nan"), dtype=M.dtype, device=M.device)], 1)
    return q.contiguous()


def kernel_inputs(st: Dict[str, torch.Tensor], layout: OscLayout, packed_M: bool = False,
                  full6_J: bool = False, with_vel: bool = False, qM: bool = False) -> Dict[str, torch.Tensor]:
    """Select the fields `BatchedOSC.step` consumes from a `synth_batch` dict."""
    keep = {
        "J": st["J6"] if full6_J else st["J"],
        "dq": st["dq"], "ee_xyz": st["ee_xyz"], "ee_quat": st["ee_quat"],
        "target_xyz": st["target_xyz"], "target_quat": st["target_quat"],
    }
    if qM:
        keep["qM"] = sparse_qM(st["M"], layout.joint_parent)
    else:
        keep["M"] = pack_lower(st["M"]) if packed_M else st["M"]
    if layout.use_g:
        keep["bias"] = st["bias"]
    if "max_vel" in st:
        keep["max_vel"] = st["max_vel"]
    if layout.admittance:
        keep["ft_xmat"] = st["ft_xmat"]
        keep["ft_raw"] = st["ft_raw"]
    if with_vel and "target_vel" in st:
        keep["target_vel"] = st["target_vel"]
    return keep


def oracle_inputs(st: Dict[str, torch.Tensor], layout: OscLayout) -> Dict[str, np.ndarray]:
    """The same batch in the field names oracle/osc_numpy.py expects (host numpy)."""
    B, D = st["dq"].shape[0], layout.D
    g = lambda k: st[k].detach().cpu().numpy()
    mv = g("max_vel") if "max_vel" in st else np.broadcast_to(
        np.array([list(d.max_vel) for d in layout.devices]), (B, D, 2)).copy()
    return {
        "M": g("M"), "J": g("J6"), "dq": g("dq"), "bias": g("bias"),
        "ee_xyz": g("ee_xyz"), "ee_quat": g("ee_quat"),
        "ft_xmat": g("ft_xmat"), "ft_raw": g("ft_raw"),
        "tgt_xyz": g("target_xyz"), "tgt_quat": g("target_quat"),
        "tgt_vel": g("target_vel") if "target_vel" in st else np.zeros((B, D, 6)),
        "max_vel": mv,
    }
