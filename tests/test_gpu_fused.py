"""GPU: the fused state provider + OSC step (`irlosc_step_fused`, osc_fused.cuh) against the oracle.

Inputs are joint states only; the oracle gets M / J / qfrc_bias / EE poses from the rigid-body
model in `dual_ur5.py` (what the reference reads from MuJoCo) for the same (q, dq).  Same
tolerance as tests/test_gpu_parity.py.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6


def _torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch


def _rel(got, want):
    scale = np.abs(want).max(axis=1, keepdims=True)
    return (np.abs(got - want) / scale).max(axis=1)


def _setup(scenario, B, seed, device="cuda:0"):
    from irl_control_b200.engine import BatchedOSC
    from irl_control_b200.synthetic import fused_inputs, scenario_model, synth_batch
    layout, model = scenario_model(scenario)
    st = synth_batch(layout, B, seed=seed, device=device, insertion_schedule=(scenario == "insertion"))
    eng = BatchedOSC(layout, device=0)
    eng.set_model(model)
    return layout, eng, st, fused_inputs(st, layout)


@pytest.mark.parametrize("scenario,B", [("gain_test", 4096), ("admit_test", 8192), ("insertion", 16384),
                                        ("worst_case", 2048), ("gain_test", 1), ("gain_test", 255),
                                        ("admit_test", 257), ("gain_test", 40000)])
def test_fused_step_matches_oracle(scenario, B):
    torch = _torch()
    from irl_control_b200 import _native
    from irl_control_b200.synthetic import oracle_inputs
    from oracle import osc_numpy
    layout, eng, st, fin = _setup(scenario, B, seed=21)
    out = eng.step_fused(fin, want_u_all=True, want_ee=True)
    torch.cuda.synchronize()
    assert eng.last_kernel.startswith("osc_step_fused")
    u_all, ctrl, status = (out[k].cpu().numpy() for k in ("u_all", "ctrl", "status"))
    n_chk = min(B, 1536)                                   # the oracle is a Python loop
    idx = np.unique(np.concatenate([np.arange(min(B, 64)), np.linspace(0, B - 1, n_chk).astype(int)]))
    # every instance that went through the eigen fix-up is checked too
    hard = np.nonzero(status & _native.ST_EIGEN)[0][:256]
    idx = np.unique(np.concatenate([idx, hard]))
    ref = osc_numpy.osc_batch(layout.as_dict(), oracle_inputs(st, layout), idx=idx)
    rel = _rel(u_all[idx], ref["u_all"])
    assert rel.max() < REL_TOL, (rel.max(), int(idx[np.argmax(rel)]))
    assert np.array_equal((status[idx] & _native.ST_PINV) != 0, ref["pinv"])
    want = np.concatenate([ref["u_all"][:, list(d.actuator_trnids)] for d in layout.devices], axis=1)
    assert _rel(ctrl[idx], want).max() < REL_TOL
    assert np.isfinite(ctrl).all() and np.isfinite(u_all).all()
    assert np.abs(out["ee_xyz"].cpu().numpy() - st["ee_xyz"].cpu().numpy()).max() < 1e-12
    assert ((status & _native.ST_EIGEN) != 0).mean() < 2e-4        # the warp eigen-solver is the rare exception
    print("fused %s B=%d: max rel %.2e, eigen fix-ups %d" % (scenario, B, rel.max(), int((status & _native.ST_EIGEN != 0).sum())))


def test_fused_equals_resident_step():
    """Same instances through the two product paths: state given (irlosc_step) vs state computed
    on the GPU (irlosc_step_fused)."""
    torch = _torch()
    from irl_control_b200.synthetic import kernel_inputs
    layout, eng, st, fin = _setup("gain_test", 8192, seed=4)
    a = eng.step(kernel_inputs(st, layout, packed_M=True), want_u_all=True)
    b = eng.step_fused(fin, want_u_all=True)
    torch.cuda.synchronize()
    rel = _rel(b["u_all"].cpu().numpy(), a["u_all"].cpu().numpy())
    assert rel.max() < REL_TOL, rel.max()


def test_fused_host_buffers_and_repeatability():
    torch = _torch()
    from irl_control_b200.engine import pinned_empty
    layout, eng, st, fin = _setup("admit_test", 70000, seed=8)       # > 2 chunks of the host pipeline
    dev_out = eng.step_fused(fin, want_u_all=True)
    torch.cuda.synchronize()
    host_in = {}
    for k, v in fin.items():
        buf = pinned_empty(tuple(v.shape))
        buf[...] = v.cpu().numpy()
        host_in[k] = buf
    h1 = eng.step_fused_host(host_in, want_u_all=True)
    h2 = eng.step_fused_host(host_in, want_u_all=True)
    for k in ("ctrl", "u_all", "status"):
        assert np.array_equal(h1[k], h2[k]), k
        assert np.array_equal(h1[k], dev_out[k].cpu().numpy()), k


def test_fused_velocity_branch_and_index_error():
    torch = _torch()
    from irl_control_b200 import _native
    from irl_control_b200.synthetic import oracle_inputs
    from oracle import osc_numpy
    B = 512
    rng = np.random.default_rng(1)
    for scenario in ("admit_test", "gain_test"):
        layout, eng, st, fin = _setup(scenario, B, seed=13)
        tv = rng.normal(0.0, 0.2, size=(B, layout.D, 6))
        tv[: B // 2, :, 0] = 0.0
        st["target_vel"] = torch.from_numpy(tv).to("cuda:0")
        fin["target_vel"] = st["target_vel"]
        out = eng.step_fused(fin, want_u_all=True)
        torch.cuda.synchronize()
        u_all, status = out["u_all"].cpu().numpy(), out["status"].cpu().numpy()
        ob = oracle_inputs(st, layout)
        for i in list(range(0, 48)) + list(range(B // 2, B // 2 + 48)):
            try:
                ref = osc_numpy.osc_step(layout.as_dict(), {k: v[i] for k, v in ob.items()})
            except IndexError:
                assert status[i] & _native.ST_DX_RANGE and np.isnan(u_all[i]).all()
                continue
            assert _rel(u_all[i:i + 1], ref["u_all"][None]).max() < REL_TOL
            assert bool(status[i] & _native.ST_VEL_BRANCH) == bool(np.any(ref["vel_branch"]))


def test_fused_requires_model_and_topology():
    _torch()
    from irl_control_b200 import _native
    from irl_control_b200.engine import BatchedOSC
    from irl_control_b200.synthetic import fused_inputs, scenario_model, synth_batch
    layout, model = scenario_model("gain_test")
    eng = BatchedOSC(layout, device=0)
    st = synth_batch(layout, 8, seed=0, device="cuda:0")
    with pytest.raises(_native.OscError, match="irlosc_set_model"):
        eng.step_fused(fused_inputs(st, layout))
    model.joint[3].parent = 0
    with pytest.raises(_native.OscError, match="parent"):
        eng.set_model(model)


def test_fused_size_independent_properties_at_full_batch():
    """B = 65 536: determinism, instance-permutation equivariance, split invariance (the launch picks a
    different warp count for the half batch), and the action-free sequence of two steps is stateless."""
    torch = _torch()
    layout, eng, st, fin = _setup("gain_test", 65536, seed=2)
    a = {k: v.clone() for k, v in eng.step_fused(fin, want_u_all=True).items()}
    b = eng.step_fused(fin, want_u_all=True)
    assert torch.equal(a["ctrl"], b["ctrl"]) and torch.equal(a["status"], b["status"])
    perm = torch.randperm(65536, device="cuda:0", generator=torch.Generator(device="cuda:0").manual_seed(3))
    c = eng.step_fused({k: v[perm].contiguous() for k, v in fin.items()}, want_u_all=True)
    assert torch.equal(c["ctrl"], a["ctrl"][perm]) and torch.equal(c["u_all"], a["u_all"][perm])
    d = eng.step_fused({k: v[30000:].contiguous() for k, v in fin.items()})
    assert torch.equal(d["ctrl"], a["ctrl"][30000:])
    cols = [j for dl in layout.devices for j in dl.actuator_trnids]
    assert torch.equal(a["ctrl"], a["u_all"][:, cols])            # packing (osc.py:203-208)
    assert torch.isfinite(a["ctrl"]).all()
