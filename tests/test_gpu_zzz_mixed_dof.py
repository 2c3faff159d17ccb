"""GPU: DoF masks no shipped YAML has (right arm xyz + b, g; left arm xyz + a; base yaw: 5 + 4 + 1 task rows).  No
specialised kernel serves unequal arm row counts, so every selector that does not force one must end in the generic
kernel, and its output must equal the reference's goldens (`mixed_dof_s15`, `mixed_dof_vel_s16`).  Written after round
1's GPU budget was spent; the oracle is checked against the same goldens on the CPU (tests/test_oracle.py)."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES_FALLBACK, GOLDEN_CASES_GENERIC, load_golden
from test_gpu_parity import REL_TOL, _golden_state, _layout_from_dict, _rel_err, _torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kernel,topology", [(1, False), (0, False), (0, True)])
@pytest.mark.parametrize("packed_M,full6_J", [(True, False), (False, True)])
@pytest.mark.parametrize("case", GOLDEN_CASES_GENERIC)
def test_generic_kernel_serves_unshipped_dof_masks(case, packed_M, full6_J, kernel, topology):
    torch = _torch()
    from irl_control_b200 import _native
    from irl_control_b200.engine import BatchedOSC
    g, ld = load_golden(case)
    layout = _layout_from_dict(ld, topology=topology, check=False)
    assert layout.k == 10
    eng = BatchedOSC(layout, device=0)
    eng.set_kernel(kernel)
    out = eng.step(_golden_state(g, layout, torch, packed_M, full6_J), want_u_all=True)
    torch.cuda.synchronize()
    assert eng.last_kernel == "osc_step_generic", eng.last_kernel
    ctrl, u_all, status = (out[k].cpu().numpy() for k in ("ctrl", "u_all", "status"))
    assert not g["index_error"].any()
    assert np.array_equal((status & _native.ST_PINV) != 0, g["pinv"])
    e_u = _rel_err(u_all, g["u_all"])
    e_c = np.abs(ctrl - g["ctrl"]).max(axis=1) / np.abs(g["u_all"]).max(axis=1)
    assert e_u.max() < REL_TOL and e_c.max() < REL_TOL
    vel = (np.asarray(g["target_vel"]) != 0).all(axis=-1).any(axis=-1)
    assert np.array_equal((status & _native.ST_VEL_BRANCH) != 0, vel)


@pytest.mark.parametrize("case", GOLDEN_CASES_GENERIC)
def test_generate_through_the_host_classes_with_unshipped_dof_masks(case):
    """Same goldens through `Device / Robot / OSC.generate` (tests/test_dropin_golden.py's body, real engine)."""
    _torch()
    from test_dropin_golden import generate_on_golden
    generate_on_golden(case)


@pytest.mark.parametrize("packed_M", [True, False])
def test_auto_dispatch_falls_back_when_the_copy_plan_does_not_fit(packed_M):
    """admittance=True with the base among the targets: the streaming kernel's copy plan would need more chunks than it
    holds, so `irlosc_step` must take a record-staging kernel (not fail) and reproduce the reference's golden; asking
    for the streaming kernel explicitly is refused."""
    torch = _torch()
    from irl_control_b200 import _native
    from irl_control_b200.engine import BatchedOSC
    g, ld = load_golden(GOLDEN_CASES_FALLBACK[0])
    layout = _layout_from_dict(ld, topology=True, check=False)
    eng = BatchedOSC(layout, device=0)
    st = _golden_state(g, layout, torch, packed_M, False)
    out = eng.step(st, want_u_all=True)
    torch.cuda.synchronize()
    assert not eng.last_kernel.startswith("osc_step_stream"), eng.last_kernel
    u_all, status = out["u_all"].cpu().numpy(), out["status"].cpu().numpy()
    assert np.array_equal((status & _native.ST_PINV) != 0, g["pinv"])
    assert _rel_err(u_all, g["u_all"]).max() < REL_TOL
    eng.set_kernel(9)
    with pytest.raises(_native.OscError):
        eng.step(st)


@pytest.mark.parametrize("case", GOLDEN_CASES_FALLBACK)
def test_generate_through_the_host_classes_with_admittance_and_base(case):
    _torch()
    from test_dropin_golden import generate_on_golden
    generate_on_golden(case)
