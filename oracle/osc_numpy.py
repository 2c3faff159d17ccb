"""TEST INFRASTRUCTURE ONLY - float64 numpy restatement of the reference OSC step.

Clean-room restatement (no code copied) of what one call of
`OSC.generate(targets)` computes for ONE robot instance, following the
reference statement by statement so that every quirk survives:

    osc.py:35-39    gain precompute                    -> `_gains`
    osc.py:41-68    __Mx / __svd_solve                 -> `task_space_inertia`, `svd_inverse`
    osc.py:70-99    __limit_vel                        -> `limit_vel`
    osc.py:101-118  calc_error                         -> `calc_error`
    osc.py:132-210  generate (assembly, per-device loop, J^T Mx, bias,
                    null space, packing)              -> `osc_step`
    robot.py:44-58  J_idxs in SUB-DEVICE order (N3)    -> layout["devices"][d]["dx_idx"]
    device.py:135-170 F/T rotation into the world      -> `rotate_ft`

The same LAPACK entry points as the reference are used (`np.linalg.svd`,
`det`, `pinv`), so branch decisions agree with the reference bit for bit on
identical inputs.  Pinned against the real reference by tests/test_oracle.py
through tests/golden/*.npz (generated with oracle/ref_harness.py).  The
quaternion helpers come from oracle/t3d.py (parity unpinned, see there).

Layout dictionary (what `Device`/`Robot`/`OSC` constructors boil down to):

    layout = {
      "n": 25,                          # Robot.num_joints_total (robot.py:32)
      "use_g": True, "admittance": False,
      "nullspace_kv": 10.0 | None,
      "devices": [                      # TARGET order (osc.py:137,156)
        {"name", "ctrlr_dof": bool[6], "joint_ids_all": int[], "actuator_trnids": int[],
         "ctrl_idxs": int[], "kp", "kv", "ko", "k": [3], "d": [3],
         "has_max_vel": bool, "dx_idx": int[k_dev]},   # J_idxs[name] (robot.py:54)
      ]}

Instance dictionary `st` (all float64):
    M (n,n) ; J (D,6,n) full [jacp;jacr] per target device ; dq (n) ; bias (n) ;
    ee_xyz (D,3) ; ee_quat (D,4) ; ft (D,6) world-frame force|torque, or
    ft_xmat (D,9) + ft_raw (D,6) sensor-frame (rotated here, device.py:135-170) ;
    tgt_xyz (D,3) ; tgt_quat (D,4) ; tgt_vel (D,6) ; max_vel (D,2)
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from . import t3d

DET_THRESHOLD = 1e-4          # osc.py:51
PINV_RCOND = DET_THRESHOLD * 0.1   # osc.py:55


def svd_inverse(A: np.ndarray) -> np.ndarray:
    """osc.py:59-68 - inverse through the SVD, no singular-value cutoff."""
    u, s, vt = np.linalg.svd(A)
    return np.dot(vt.transpose(), np.dot(np.diag(s ** -1), u.transpose()))


def task_space_inertia(J: np.ndarray, M: np.ndarray) -> Tuple[np.ndarray, np.ndarray, bool]:
    """osc.py:41-56.  Returns (Mx, M_inv, pinv_branch_taken)."""
    M_inv = svd_inverse(M)
    Mx_inv = np.dot(J, np.dot(M_inv, J.T))
    if abs(np.linalg.det(Mx_inv)) >= DET_THRESHOLD:
        return svd_inverse(Mx_inv), M_inv, False
    return np.linalg.pinv(Mx_inv, rcond=PINV_RCOND), M_inv, True


def _gains(dev: Dict) -> Tuple[np.ndarray, np.ndarray]:
    """osc.py:36-39."""
    if dev.get("gain_vectors") is not None:          # what OSC.__init__ stored (osc.py:37-39); kp / kv / ko may have changed since
        return np.array(dev["gain_vectors"][0], dtype=np.float64), np.array(dev["gain_vectors"][1], dtype=np.float64)
    task_space_gains = np.array([dev["kp"]] * 3 + [dev["ko"]] * 3, dtype=np.float64)
    lamb = task_space_gains / dev["kv"]
    return task_space_gains, lamb


def calc_error(dev: Dict, ee_xyz, ee_quat, tgt_xyz, tgt_quat) -> np.ndarray:
    """osc.py:101-118 - note the result is NOT masked by ctrlr_dof."""
    dof = np.asarray(dev["ctrlr_dof"], dtype=bool)
    u_task = np.zeros(6)
    if np.sum(dof[:3]) > 0:
        u_task[:3] = np.asarray(ee_xyz) - np.asarray(tgt_xyz)
    if np.sum(dof[3:]) > 0:
        q_d = t3d.normalized_vector(tgt_quat)
        q_r = np.array(t3d.qmult(q_d, t3d.qconjugate(ee_quat)))
        u_task[3:] = t3d.quat2euler(t3d.qconjugate(q_r))
    return u_task


def limit_vel(dev: Dict, u_task: np.ndarray, max_vel) -> np.ndarray:
    """osc.py:70-99 (the `max_vel is None` raise is unreachable from generate)."""
    kv, kp, ko = dev["kv"], dev["kp"], dev["ko"]
    _, lamb = _gains(dev)
    scale = np.ones(6)
    norm_xyz = np.linalg.norm(u_task[:3])
    sat_xyz = max_vel[0] / kp * kv
    if norm_xyz > sat_xyz:
        scale[:3] *= sat_xyz / norm_xyz
    norm_abg = np.linalg.norm(u_task[3:])
    sat_abg = max_vel[1] / ko * kv
    if norm_abg > sat_abg:
        scale[3:] *= sat_abg / norm_abg
    return kv * scale * lamb * u_task


def rotate_ft(xmat: np.ndarray, raw6: np.ndarray) -> np.ndarray:
    """device.py:135-170 - force = R @ sensordata[f:f+3], torque = R @ sensordata[t:t+3]."""
    R = np.asarray(xmat, dtype=np.float64).reshape(3, 3)
    return np.concatenate([R @ raw6[:3], R @ raw6[3:]])


def osc_step(layout: Dict, st: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """One `OSC.generate` for one instance.  Returns u_all (n), ctrl (packed forces in
    target order), forces (list per target), pinv (bool), vel_branch (bool per device)."""
    n = layout["n"]
    devs: List[Dict] = layout["devices"]
    M = np.asarray(st["M"], dtype=np.float64)
    dq = np.asarray(st["dq"], dtype=np.float64)

    # osc.py:136-138 - stack the row-masked Jacobians in target order
    rows = []
    for d, dev in enumerate(devs):
        dof = np.asarray(dev["ctrlr_dof"], dtype=bool)
        rows.append(np.asarray(st["J"][d], dtype=np.float64)[dof])   # device.py:132
    J = np.vstack(rows)

    Mx, M_inv, pinv_taken = task_space_inertia(J, M)

    dx = np.dot(J, dq)                 # osc.py:150
    uv_all = np.dot(M, dq)             # osc.py:151
    u_all = np.zeros(n)
    u_task_all = np.array([])
    ext_f = np.array([])
    vel_branch = []

    for d, dev in enumerate(devs):
        dof = np.asarray(dev["ctrlr_dof"], dtype=bool)
        u_task = calc_error(dev, st["ee_xyz"][d], st["ee_quat"][d], st["tgt_xyz"][d], st["tgt_quat"][d])
        stiffness = np.array(list(dev["k"]) + [1] * 3, dtype=np.float64)
        damping = np.array(list(dev["d"]) + [1] * 3, dtype=np.float64)
        if dev["has_max_vel"]:
            u_task = limit_vel(dev, u_task, st["max_vel"][d])
            u_task = u_task * stiffness
        else:
            gains, _ = _gains(dev)
            u_task = u_task * (gains * stiffness)

        kv = dev["kv"]
        target_vel = np.asarray(st["tgt_vel"][d], dtype=np.float64)
        ids = np.asarray(dev["joint_ids_all"], dtype=np.int64)
        if np.all(target_vel) == 0:        # osc.py:173 (N4): True unless ALL six are non-zero
            u_all[ids] = -1 * kv * uv_all[ids]
            vel_branch.append(False)
        else:
            diff = dx[np.asarray(dev["dx_idx"], dtype=np.int64)] - target_vel[dof]   # may raise IndexError (N3)
            u_task[dof] += kv * diff * damping[dof]
            vel_branch.append(True)

        if "ft" in st:
            wrench = np.asarray(st["ft"][d], dtype=np.float64)
        else:
            wrench = rotate_ft(st["ft_xmat"][d], np.asarray(st["ft_raw"][d], dtype=np.float64))
        ext_f = np.append(ext_f, wrench[dof])
        u_task_all = np.append(u_task_all, u_task[dof])

    if layout["admittance"]:
        u_all -= np.dot(J.T, np.dot(Mx, u_task_all + ext_f))
    else:
        u_all -= np.dot(J.T, np.dot(Mx, u_task_all))

    if layout["use_g"]:
        u_all += np.asarray(st["bias"], dtype=np.float64)

    if layout.get("nullspace_kv") is not None:
        u_null = np.dot(M, -layout["nullspace_kv"] * dq)
        Jbar = np.dot(M_inv, np.dot(J.T, Mx))
        null_filter = np.eye(n) - np.dot(J.T, Jbar.T)
        u_all += np.dot(null_filter, u_null)

    forces = [u_all[np.asarray(dev["actuator_trnids"], dtype=np.int64)] for dev in devs]
    return {
        "u_all": u_all,
        "forces": forces,
        "ctrl": np.concatenate(forces) if forces else np.zeros(0),
        "pinv": pinv_taken,
        "vel_branch": np.array(vel_branch, dtype=bool),
        "det": float(np.linalg.det(np.dot(J, np.dot(M_inv, J.T)))),
    }


def osc_batch(layout: Dict, batch: Dict[str, np.ndarray], idx=None) -> Dict[str, np.ndarray]:
    """Plain loop over instances of a batch dict (leading axis B on every array)."""
    B = batch["M"].shape[0]
    sel = range(B) if idx is None else idx
    outs = [osc_step(layout, {k: v[i] for k, v in batch.items()}) for i in sel]
    return {
        "u_all": np.stack([o["u_all"] for o in outs]),
        "ctrl": np.stack([o["ctrl"] for o in outs]),
        "pinv": np.array([o["pinv"] for o in outs], dtype=bool),
        "det": np.array([o["det"] for o in outs]),
        "vel_branch": np.stack([o["vel_branch"] for o in outs]),
    }


def full_from_qM(qM, dof_parentid):
    """Dense n x n inertia from MuJoCo's sparse `qM`: restatement of `mj_fullM`, which the reference calls at
    robot.py:69 (`mujoco_py.cymj._mj_fullM(model, M_vec, sim.data.qM)`).  MuJoCo is a third-party dependency
    absent from /root/reference and from this image (parity unpinned); published algorithm (MuJoCo 2.x
    engine_util_misc.c): `adr = dof_Madr[i]; for (j = i; j >= 0; j = dof_parentid[j]) dst[i][j] = dst[j][i] =
    qM[adr++]`, with dof_Madr the running sum of the dofs' depths."""
    n = len(dof_parentid)
    M = np.zeros((n, n))
    adr = 0
    for i in range(n):
        j = i
        while j >= 0:
            M[i, j] = M[j, i] = qM[adr]
            adr += 1
            j = dof_parentid[j]
    return M
