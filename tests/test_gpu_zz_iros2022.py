"""GPU: SURVEY 8 (f4) - the iros2022 configuration (robot_configs/iros2022.yaml gains and max_vel, devices and
order of action_sequence_configs/iros2022_task.yaml:1-4: base, ur5left, ur5right; k = 13) through the C ABI.

The same parametrised test also runs the other late goldens of `GOLDEN_CASES_F4`: `device.max_vel = None` on some
devices (osc.py:163-168, the un-limited gain branch; layouts with `has_max_vel` False).

Written after round 1's GPU budget was spent: the goldens (reference outputs) and the host build of the
default kernel are checked in the CPU suite (tests/test_oracle.py, tests/test_stream_host.py,
tests/test_fused_host.py); this file is their first run on a GPU, kept apart from tests/test_gpu_parity.py
and last in collection order for that reason.
"""
import numpy as np
import pytest

from conftest import GOLDEN_CASES_F4, load_golden
from test_gpu_parity import REL_TOL, _golden_state, _layout_from_dict, _rel_err, _torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kernel,topology", [(0, True), (9, True), (1, False), (0, False), (2, True)])
@pytest.mark.parametrize("packed_M,full6_J", [(True, False), (False, False), (True, True)])
@pytest.mark.parametrize("case", GOLDEN_CASES_F4)
def test_cuda_matches_reference_golden_iros2022(case, packed_M, full6_J, kernel, topology):
    torch = _torch()
    from irl_control_b200 import _native
    from irl_control_b200.engine import BatchedOSC
    g, ld = load_golden(case)
    if kernel == 2 and full6_J:
        pytest.skip("the tree-sparse kernel stages the row-stacked Jacobian layout only")
    # check_topology makes the kernels that read every entry eligible only, so not with the streaming kernel
    layout = _layout_from_dict(ld, topology=topology, check=(topology and kernel == 2))
    eng = BatchedOSC(layout, device=0)
    eng.set_kernel(kernel)
    out = eng.step(_golden_state(g, layout, torch, packed_M, full6_J), want_u_all=True)
    torch.cuda.synchronize()
    ctrl, u_all, status = (out[k].cpu().numpy() for k in ("ctrl", "u_all", "status"))
    assert not g["index_error"].any()
    assert np.array_equal((status & _native.ST_PINV) != 0, g["pinv"])
    assert not np.any(status & (_native.ST_M_NOT_PD | _native.ST_DX_RANGE | _native.ST_SPARSITY))
    e_u = _rel_err(u_all, g["u_all"])
    e_c = np.abs(ctrl - g["ctrl"]).max(axis=1) / np.abs(g["u_all"]).max(axis=1)
    print("%s kernel=%s worst rel err u_all %.2e ctrl %.2e" % (case, eng.last_kernel, e_u.max(), e_c.max()))
    assert e_u.max() < REL_TOL and e_c.max() < REL_TOL
    vel = (np.asarray(g["target_vel"]) != 0).all(axis=-1).any(axis=-1)
    assert np.array_equal((status & _native.ST_VEL_BRANCH) != 0, vel)


@pytest.mark.parametrize("B", [4096, 65536])
def test_iros2022_batch_matches_oracle_and_fused_step(B):
    """Synthetic iros2022 batch: `M, J` step and fused (q, dq) step against the oracle on a strided subset."""
    torch = _torch()
    from irl_control_b200 import _native
    from irl_control_b200.engine import BatchedOSC
    from irl_control_b200.synthetic import fused_inputs, kernel_inputs, oracle_inputs, scenario_model, synth_batch
    from oracle import osc_numpy
    layout, model = scenario_model("iros2022")
    st = synth_batch(layout, B, seed=B + 5, device="cuda:0")
    eng = BatchedOSC(layout, device=0)
    out = eng.step(kernel_inputs(st, layout), want_u_all=True)
    torch.cuda.synchronize()
    name = eng.last_kernel
    u_all, status = out["u_all"].cpu().numpy(), out["status"].cpu().numpy()
    assert np.isfinite(u_all).all() and not np.any(status & (_native.ST_M_NOT_PD | _native.ST_DX_RANGE))
    idx = np.arange(0, B, max(1, B // 300))
    ref = osc_numpy.osc_batch(layout.as_dict(), oracle_inputs(st, layout), idx=idx)
    agree = ((status[idx] & _native.ST_PINV) != 0) == ref["pinv"]
    near = np.abs(np.abs(ref["det"]) - 1e-4) < 1e-9
    assert np.all(agree | near)
    err = _rel_err(u_all[idx], ref["u_all"])
    assert err[agree].max() < REL_TOL
    eng.set_model(model)
    fo = eng.step_fused(fused_inputs(st, layout), want_u_all=True)
    torch.cuda.synchronize()
    ferr = _rel_err(fo["u_all"].cpu().numpy()[idx], ref["u_all"])
    print("iros2022 B=%d: %s worst %.2e; %s worst %.2e; pinv share %.3f" % (
        B, name, err[agree].max(), eng.last_kernel, ferr[agree].max(), ref["pinv"].mean()))
    assert ferr[agree].max() < REL_TOL
