"""CPU: the streaming step (`csrc/osc_stream.cuh` copy plan + per-instance elimination, `osc_tail.cuh`)
compiled for the host (tests/host_fused, test infrastructure) against the reference's golden outputs.

Same assertions as `tests/test_gpu_parity.py::test_streaming_kernel_matches_reference_golden`, so the
arithmetic, the gather plan of every `M` / `J` layout and the status flags of the default kernel of the
6-row configurations (admit_test, insertion, k = 13, iros2022) are checked without a GPU; what only the
GPU suite sees is the cp.async staging and the fix-up kernel's warp code.
"""
import numpy as np
import pytest

import fused_host
from conftest import GOLDEN_CASES, GOLDEN_CASES_F4, load_golden
from irl_control_b200 import _native
from irl_control_b200.layout import DeviceLayout, OscLayout, qm_index, qm_size
from oracle import osc_numpy

REL_TOL = 1e-6
DUAL_UR5_PARENT = (-1, 0, 1, 2, 3, 4, 5, 6, 7, 6, 6, 10, 6, 0, 13, 14, 15, 16, 17, 18, 19, 18, 18, 22, 18)
EE_JOINT = {"base": 0, "ur5right": 6, "ur5left": 18}


def _layout(ld):
    devs = tuple(DeviceLayout(name=d["name"], ctrlr_dof=tuple(d["ctrlr_dof"]), joint_ids_all=tuple(d["joint_ids_all"]),
                              actuator_trnids=tuple(d["actuator_trnids"]), ctrl_idxs=tuple(d["ctrl_idxs"]),
                              dx_idx=tuple(d["dx_idx"]), has_max_vel=d["has_max_vel"], max_vel=tuple(d["max_vel"]),
                              kp=d["kp"], kv=d["kv"], ko=d["ko"], k=tuple(d["k"]), d=tuple(d["d"]),
                              ee_joint=EE_JOINT[d["name"]]) for d in ld["devices"])
    return OscLayout(n=ld["n"], devices=devs, use_g=ld["use_g"], admittance=ld["admittance"],
                     nullspace_kv=ld["nullspace_kv"], joint_parent=DUAL_UR5_PARENT, check_topology=False)


def _state(g, layout, packed_M, full6_J, qM_pad=None):
    rows = [(d, c) for d, dl in enumerate(layout.devices) for c in range(6) if dl.ctrlr_dof[c]]
    st = {k: np.array(g[k]) for k in ("M", "dq", "bias", "ee_xyz", "ee_quat", "target_xyz", "target_quat", "max_vel")}
    st["J"] = np.array(g["J6"]) if full6_J else np.stack([g["J6"][:, d, c] for d, c in rows], 1)
    if qM_pad is not None:
        rows, cols = qm_index(layout.joint_parent)
        q = st.pop("M")[:, rows, cols]
        st["qM"] = np.ascontiguousarray(np.concatenate([q, np.full((q.shape[0], qM_pad), np.nan)], 1))
    elif packed_M:
        n = layout.n
        il = np.tril_indices(n)
        st["M"] = np.ascontiguousarray(st["M"][:, il[0], il[1]])
    if layout.admittance:
        st["ft_xmat"], st["ft_raw"] = np.array(g["ft_xmat"]), np.array(g["ft_raw"])
    if np.any(g["target_vel"] != 0):
        st["target_vel"] = np.array(g["target_vel"])
    return st


@pytest.mark.parametrize("packed_M,full6_J", [(False, False), (True, True), (True, False)])
@pytest.mark.parametrize("case", GOLDEN_CASES + GOLDEN_CASES_F4)
def test_host_build_of_the_streaming_step_matches_reference_golden(case, packed_M, full6_J):
    g, ld = load_golden(case)
    layout = _layout(ld)
    out = fused_host.run_stream(layout, _state(g, layout, packed_M, full6_J))
    ctrl, u_all, status = out["ctrl"], out["u_all"], out["status"]
    bad = np.array(g["index_error"])
    assert np.all((status[bad] & _native.ST_DX_RANGE) != 0) and np.all(np.isnan(ctrl[bad]))
    ok = ~bad
    if ok.any():
        assert np.array_equal((status[ok] & _native.ST_PINV) != 0, g["pinv"][ok])
        assert not np.any(status[ok] & (_native.ST_M_NOT_PD | _native.ST_DX_RANGE | _native.ST_SPARSITY))
        scale = np.abs(g["u_all"][ok]).max(axis=1)
        e_u = np.abs(u_all[ok] - g["u_all"][ok]).max(axis=1) / scale
        e_c = np.abs(ctrl[ok] - g["ctrl"][ok]).max(axis=1) / scale
        assert e_u.max() < REL_TOL and e_c.max() < REL_TOL, (case, e_u.max(), e_c.max())
        vel = (np.asarray(g["target_vel"]) != 0).all(axis=-1).any(axis=-1)
        assert np.array_equal((status[ok] & _native.ST_VEL_BRANCH) != 0, vel[ok])


def test_qm_index_is_the_mj_fullM_walk():
    """`layout.qm_index` (product) and `oracle.full_from_qM` (restatement of mj_fullM, robot.py:69) are inverse
    to each other on the DualUR5 tree; nM = 155 (155 of 325 lower entries are structural)."""
    g, ld = load_golden("gain_test_s0")
    rows, cols = qm_index(DUAL_UR5_PARENT)
    assert qm_size(DUAL_UR5_PARENT) == 155 and len(rows) == 155
    assert (rows[0], cols[0]) == (0, 0) and list(zip(rows[1:3], cols[1:3])) == [(1, 1), (1, 0)]
    for i in range(4):
        M = np.array(g["M"][i])
        back = osc_numpy.full_from_qM(M[rows, cols], DUAL_UR5_PARENT)
        assert np.array_equal(back, M)          # every entry outside the walk is an exact zero of the tree


@pytest.mark.parametrize("full6_J,pad", [(False, 0), (True, 42)])
@pytest.mark.parametrize("case", GOLDEN_CASES + GOLDEN_CASES_F4)
def test_host_build_of_the_streaming_step_reads_mujoco_sparse_qM(case, full6_J, pad):
    """IRLOSC_M_QM: the step fed with what `sim.data.qM` holds (robot.py:69 expands it with mj_fullM) - tight,
    and padded like a scene whose free bodies add entries after the robot's - reproduces the goldens."""
    g, ld = load_golden(case)
    layout = _layout(ld)
    dense = fused_host.run_stream(layout, _state(g, layout, False, full6_J))
    out = fused_host.run_stream(layout, _state(g, layout, False, full6_J, qM_pad=pad))
    ok = ~np.array(g["index_error"])
    assert np.array_equal(out["status"], dense["status"])
    # same entries, same arithmetic: bit-identical to the dense-M run
    assert np.array_equal(out["u_all"][ok], dense["u_all"][ok]) and np.array_equal(out["ctrl"][ok], dense["ctrl"][ok])
    if ok.any():
        scale = np.abs(g["u_all"][ok]).max(axis=1)
        assert (np.abs(out["u_all"][ok] - g["u_all"][ok]).max(axis=1) / scale).max() < REL_TOL
        assert (np.abs(out["ctrl"][ok] - g["ctrl"][ok]).max(axis=1) / scale).max() < REL_TOL
    assert out["n_chunks"] <= dense["n_chunks"]


def test_engine_accepts_qM_in_place_of_M():
    """Host-side handling of `state["qM"]` (no GPU: the object is built without `irlosc_create`)."""
    from irl_control_b200.engine import BatchedOSC
    _, ld = load_golden("gain_test_s0")
    eng = BatchedOSC.__new__(BatchedOSC)
    eng._handle = None
    eng.layout = _layout(ld)
    J = np.zeros((4, 7, 25))
    st = eng._accept_qM({"qM": np.zeros((4, 155 + 21)), "J": J})
    assert "qM" not in st and st["M"].shape == (4, 176)
    assert eng._infer_layouts(st) == (_native.M_QM, _native.J_ROWS)
    assert eng._infer_layouts({"M": np.zeros((4, 325)), "J": J}) == (_native.M_PACKED, _native.J_ROWS)
    with pytest.raises(ValueError):
        eng._accept_qM({"qM": np.zeros((4, 154)), "J": J})
    with pytest.raises(ValueError):
        eng._accept_qM({"qM": np.zeros((4, 155)), "M": np.zeros((4, 325)), "J": J})


def test_adapter_checks_the_models_qM_addressing():
    """`mujoco_adapter.sparse_inertia` on a stand-in carrying MjModel's `dof_parentid` / `dof_Madr`."""
    from types import SimpleNamespace
    from irl_control_b200.mujoco_adapter import sparse_inertia
    _, ld = load_golden("gain_test_s0")
    layout = _layout(ld)
    free = [-1, 25, 26, 27, 28, 29]                        # one free body after the robot: dofs 25..30
    parent = list(DUAL_UR5_PARENT) + free
    depth = []
    for i in range(len(parent)):
        d, j = 0, i
        while j >= 0:
            d, j = d + 1, parent[j]
        depth.append(d)
    madr = np.concatenate([[0], np.cumsum(depth)[:-1]])
    m = SimpleNamespace(dof_parentid=np.array(parent), dof_Madr=madr)
    data = SimpleNamespace(qM=np.arange(float(sum(depth))))
    assert sum(depth) == 155 + 21
    assert sparse_inertia(m, data, layout).shape == (176,)
    with pytest.raises(ValueError):                         # robot not first in the scene
        sparse_inertia(SimpleNamespace(dof_parentid=np.array(free + parent[:25]), dof_Madr=madr), data, layout)
    with pytest.raises(ValueError):
        sparse_inertia(SimpleNamespace(dof_parentid=np.array(parent), dof_Madr=madr + 1), data, layout)
    with pytest.raises(ValueError):
        sparse_inertia(m, SimpleNamespace(qM=np.zeros(100)), layout)


def test_tree_kernel_qM_offsets_are_the_mj_fullM_walk():
    """The tree-sparse kernel's qM instantiations (csrc/osc_tree.cuh) address the record as per-lane base +
    compile-time offsets; the same expressions, evaluated here, must hit the slots `layout.qm_index` assigns."""
    rows, cols = qm_index(DUAL_UR5_PARENT)
    off = {(int(r), int(c)): k for k, (r, c) in enumerate(zip(rows, cols))}
    adr = {}
    for k, r in enumerate(rows):
        adr.setdefault(int(r), k)
    assert (adr[1], adr[7], adr[13], adr[24] + 8) == (1, 28, 78, 155)          # the kernel's static_assert
    for arm in (0, 1):
        jb, qa = 1 + 12 * arm, (adr[13] if arm else adr[1])
        ccol = lambda i: 0 if i == 0 else jb + i - 1                           # noqa: E731
        for i in range(1, 7):                                                  # arm chain rows
            ro = qa + (i - 1) * (i + 2) // 2
            for j in range(i + 1):
                assert ro + (i if j == 0 else i - j) == off[(jb + i - 1, ccol(j))]
        for h in (0, 1):                                                       # gripper halves
            gb = jb + 6 + 3 * h
            for r in range(3):
                qs = 1 if r == 1 else 0
                ro = qa + 27 + 25 * h + (0, 8, 17)[r]
                assert ro == off[(gb + r, gb + r)]
                for i in range(7):
                    assert ro + 7 + qs - i == off[(gb + r, ccol(i))]
            assert qa + 27 + 25 * h + 9 == off[(gb + 1, gb)]


def test_copy_plan_limit_is_reported_for_three_devices_with_admittance():
    """admittance=True with the base among the targets needs 74 copy-plan chunks, the streaming kernel holds 72: the
    plan builder says so (auto dispatch in `irlosc_step` then takes a record-staging kernel,
    tests/test_gpu_zzz_mixed_dof.py)."""
    from conftest import GOLDEN_CASES_FALLBACK
    g, ld = load_golden(GOLDEN_CASES_FALLBACK[0])
    layout = _layout(ld)
    with pytest.raises(RuntimeError, match="more than 72 chunks"):
        fused_host.run_stream(layout, _state(g, layout, True, False))
