"""CPU: the lane step (`csrc/osc_lane.cuh`) - tile layout, pack table and the kernel's per-instance function
compiled for the host (tests/host_fused, test infrastructure) - against the reference's golden outputs.

The tile layout is what `irlosc_pack_tiles[_host]` writes and `irlosc_step_tiles` reads; here the same C++
builds it from every `M` / `J` layout of `irlosc_io`, and a numpy packer written from the exported entry table
alone (`irlosc_tile_spec`) must produce the same bytes.  What only the GPU suite sees are the coalesced loads
and the warp-cooperative finish.
"""
import numpy as np
import pytest

import fused_host
from conftest import GOLDEN_CASES, GOLDEN_CASES_F4, load_golden
from irl_control_b200 import _native
from irl_control_b200.layout import qm_index
from test_stream_host import _layout, _state

REL_TOL = 1e-6
ARR = {1: "M", 2: "J", 3: "dq", 4: "bias", 5: "ee_xyz", 6: "ee_quat", 7: "target_xyz", 8: "target_quat", 9: "max_vel",
       10: "ft_xmat", 11: "ft_raw"}


def numpy_tiles(layout, g, spec):
    """Tiles from the entry table alone: tiles[t][e][l] = entry e of instance 32 t + l (padding repeats the last)."""
    B = g["dq"].shape[0]
    rows = [(d, c) for d, dl in enumerate(layout.devices) for c in range(6) if dl.ctrlr_dof[c]]
    J = np.stack([g["J6"][:, d, c] for d, c in rows], 1)
    cols = []
    for arr, i, j in spec:
        if arr == 0:
            cols.append(np.zeros(B))
        elif arr == 1:
            cols.append(g["M"][:, i, j])
        elif arr == 2:
            cols.append(J[:, i, j])
        elif arr in (3, 4):
            cols.append(g[ARR[arr]][:, i])          # bias is copied when given (multiplied by 0 when !use_g)
        elif arr == 10:
            cols.append(g["ft_xmat"].reshape(B, -1, 9)[:, i, j])
        else:
            cols.append(g[ARR[arr]][:, i, j])
    flat = np.stack(cols, 1)                                   # [B][E]
    nt = (B + 31) // 32
    idx = np.minimum(np.arange(nt * 32), B - 1)
    return np.ascontiguousarray(flat[idx].reshape(nt, 32, -1).transpose(0, 2, 1))


@pytest.mark.parametrize("packed_M,full6_J,qM_pad", [(False, False, None), (True, True, None), (True, False, None),
                                                     (False, False, 42)])
@pytest.mark.parametrize("case", GOLDEN_CASES + GOLDEN_CASES_F4)
def test_host_build_of_the_lane_step_matches_reference_golden(case, packed_M, full6_J, qM_pad):
    g, ld = load_golden(case)
    layout = _layout(ld)
    out = fused_host.run_lane(layout, _state(g, layout, packed_M, full6_J, qM_pad))
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


@pytest.mark.parametrize("case", ["gain_test_s0", "admit_test_s1", "worst_case_s3", "gain_test_no_g_s12"])
def test_tile_layout_follows_the_exported_entry_table(case):
    """The product's packer and a numpy packer that knows nothing but `irlosc_tile_spec` write identical tiles; the
    table itself only names tree non-zeros, every M entry of MuJoCo's qM set exactly once (arm 1 repeats M[0][0] and
    dq[0], which the elimination reads per arm)."""
    g, ld = load_golden(case)
    layout = _layout(ld)
    spec, gbase = fused_host.lane_spec(layout)
    assert gbase[0] == 0 and gbase[11] == len(spec) and all(a < b for a, b in zip(gbase, gbase[1:]))
    out = fused_host.run_lane(layout, _state(g, layout, True, False))
    want = numpy_tiles(layout, {k: np.array(g[k]) for k in g.files if k != "layout_json"}, spec)
    assert out["tiles"].shape == want.shape and np.array_equal(out["tiles"], want)
    m_entries = [(i, j) for arr, i, j in spec if arr == 1]
    rows, cols = qm_index(layout.joint_parent)
    assert sorted(set(m_entries)) == sorted(zip(rows, cols))
    assert len(m_entries) == len(rows) + 1                      # M[0][0] once per arm
    kd = sum(layout.devices[0].ctrlr_dof) if layout.devices[0].name != "base" else sum(layout.devices[1].ctrlr_dof)
    n_j = sum(1 for arr, _, _ in spec if arr == 2)
    assert n_j == 2 * 7 * kd + (1 if any(d.name == "base" for d in layout.devices) else 0)


def test_velocity_term_with_non_uniform_joint_ownership():
    """`u_all[dev.joint_ids_all] = -kv * uv_all[...]` (osc.py:174) when a device's joint_ids_all is NOT a whole arm - no
    shipped YAML does that, the kernels then evaluate the coefficient per joint instead of per group - against the
    oracle, for the lane and the streaming step."""
    import dataclasses
    from irl_control_b200.synthetic import oracle_inputs, scenario_layout, synth_batch
    from oracle import osc_numpy
    layout = scenario_layout("gain_test")
    devs = []
    for d in layout.devices:
        if d.name == "ur5right":
            d = dataclasses.replace(d, joint_ids_all=tuple(j for j in d.joint_ids_all if j not in (3, 8, 12)))
        devs.append(d)
    layout = dataclasses.replace(layout, devices=tuple(devs))
    st = synth_batch(layout, 96, seed=41)
    state = {k: st[k].numpy() for k in ("M", "J", "dq", "bias", "ee_xyz", "ee_quat", "target_xyz", "target_quat", "max_vel")}
    ref = osc_numpy.osc_batch(layout.as_dict(), oracle_inputs(st, layout))
    scale = np.abs(ref["u_all"]).max(axis=1)
    for run in (fused_host.run_lane, fused_host.run_stream):
        out = run(layout, state)
        assert (np.abs(out["u_all"] - ref["u_all"]).max(axis=1) / scale).max() < REL_TOL
    # and it matters: with whole-arm ownership the excluded joints get the -kv (M dq) term
    full = osc_numpy.osc_batch(scenario_layout("gain_test").as_dict(), oracle_inputs(st, layout))
    assert np.abs(full["u_all"][:, [3, 8, 12]] - ref["u_all"][:, [3, 8, 12]]).max() > 1e-6
