"""GPU: IRLOSC_M_QM - the step fed with MuJoCo's sparse inertia `mjData.qM` (the array robot.py:69 expands
with mj_fullM) instead of a dense / packed `M`.

Written after round 1's GPU budget was spent.  The copy plan and the arithmetic are checked in the CPU suite on
the host build of the streaming step (tests/test_stream_host.py); this file is the layout's first run on a GPU
and is collected last for that reason.
"""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, GOLDEN_CASES_F4, load_golden
from test_gpu_parity import REL_TOL, _golden_state, _layout_from_dict, _rel_err, _torch

pytestmark = pytest.mark.gpu


def _with_qM(st, layout, pad):
    from irl_control_b200.synthetic import sparse_qM
    st = dict(st)
    st["qM"] = sparse_qM(st.pop("M"), layout.joint_parent, pad=pad)
    return st


@pytest.mark.parametrize("full6_J,pad", [(False, 0), (True, 42)])
@pytest.mark.parametrize("case", GOLDEN_CASES + GOLDEN_CASES_F4)
def test_streaming_kernel_reads_sparse_qM(case, full6_J, pad):
    torch = _torch()
    from irl_control_b200 import _native
    from irl_control_b200.engine import BatchedOSC
    g, ld = load_golden(case)
    layout = _layout_from_dict(ld, topology=True, check=False)
    eng = BatchedOSC(layout, device=0)
    dense_state = _golden_state(g, layout, torch, False, full6_J)
    eng.set_kernel(9)
    dense = eng.step(dense_state, want_u_all=True)
    out = eng.step(_with_qM(dense_state, layout, pad), want_u_all=True)     # kernel 9 on qM: the same copy plan entries
    torch.cuda.synchronize()
    assert eng.last_kernel.startswith("osc_step_stream"), eng.last_kernel
    # auto: 3-row arm devices with a tight qM and row-stacked J go to the tree-sparse kernel's qM instantiation,
    # everything else to the streaming kernel
    eng.set_kernel(0)
    auto = eng.step(_with_qM(dense_state, layout, pad), want_u_all=True)
    torch.cuda.synchronize()
    kd3 = all(sum(d.ctrlr_dof) == 3 for d in layout.devices if d.name != "base")
    want_tree = kd3 and not full6_J and pad == 0 and "target_vel" not in dense_state
    assert ("osc_step_tree" in eng.last_kernel and "qM" in eng.last_kernel) or not want_tree, eng.last_kernel
    ok = ~np.array(g["index_error"])
    u, u_dense = out["u_all"].cpu().numpy(), dense["u_all"].cpu().numpy()
    assert np.array_equal(out["status"].cpu().numpy(), dense["status"].cpu().numpy())
    assert np.array_equal(u[ok], u_dense[ok])           # same entries, same arithmetic
    if ok.any():
        assert _rel_err(u[ok], g["u_all"][ok]).max() < REL_TOL
        e_c = np.abs(out["ctrl"].cpu().numpy()[ok] - g["ctrl"][ok]).max(axis=1) / np.abs(g["u_all"][ok]).max(axis=1)
        assert e_c.max() < REL_TOL
        assert np.array_equal((out["status"].cpu().numpy()[ok] & _native.ST_PINV) != 0, g["pinv"][ok])
        assert _rel_err(auto["u_all"].cpu().numpy()[ok], g["u_all"][ok]).max() < REL_TOL
        assert np.array_equal((auto["status"].cpu().numpy()[ok] & _native.ST_PINV) != 0, g["pinv"][ok])


def test_qM_full_batch_host_buffers_and_refusals():
    """B = 65 536 gain_test: qM run == packed-M run bit for bit (device and host-buffer entry points); a kernel
    that cannot address qM refuses it instead of misreading it."""
    torch = _torch()
    from irl_control_b200 import _native
    from irl_control_b200.engine import BatchedOSC
    from irl_control_b200.synthetic import kernel_inputs, scenario_layout, synth_batch
    layout = scenario_layout("gain_test")
    B = 65536
    st = synth_batch(layout, B, seed=77, device="cuda:0")
    eng = BatchedOSC(layout, device=0)
    eng.set_kernel(9)
    a = eng.step(kernel_inputs(st, layout, packed_M=True), want_u_all=True)
    qin = kernel_inputs(st, layout, qM=True)
    assert tuple(qin["qM"].shape) == (B, 155)
    b = eng.step(qin, want_u_all=True)
    torch.cuda.synchronize()
    assert torch.equal(a["u_all"], b["u_all"]) and torch.equal(a["ctrl"], b["ctrl"]) and torch.equal(a["status"], b["status"])
    # auto dispatch: the tree-sparse kernel on qM (default for this layout), bit-identical to its packed-M run
    eng.set_kernel(0)
    c = eng.step(qin, want_u_all=True)
    assert "osc_step_tree" in eng.last_kernel and "qM" in eng.last_kernel, eng.last_kernel
    d = eng.step(kernel_inputs(st, layout, packed_M=True), want_u_all=True)
    assert "osc_step_tree" in eng.last_kernel and "packed" in eng.last_kernel, eng.last_kernel
    assert torch.equal(c["u_all"], d["u_all"]) and torch.equal(c["ctrl"], d["ctrl"]) and torch.equal(c["status"], d["status"])
    scale = a["u_all"].abs().amax(dim=1, keepdim=True)
    assert ((c["u_all"] - a["u_all"]).abs() / scale).max().item() < REL_TOL
    n = 3000
    host = {k: v[:n].cpu().numpy() for k, v in qin.items()}
    h = eng.step_host(host, want_u_all=True)
    assert np.array_equal(h["u_all"], c["u_all"][:n].cpu().numpy())
    eng.set_kernel(1)                                 # the generic kernel cannot address qM
    with pytest.raises(_native.OscError):
        eng.step(qin)
    eng.set_kernel(0)
    with pytest.raises(ValueError):
        eng.step(dict(qin, M=st["M"]))
