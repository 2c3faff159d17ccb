"""GPU: the two kernels on batch-interleaved tiles - lane (`csrc/osc_lane.cuh`, a thread per instance) and pair
(`csrc/osc_pair.cuh`, a lane per arm) - through the C ABI (`irlosc_pack_tiles`, `irlosc_pack_tiles_host`,
`irlosc_step_tiles`, `irlosc_step_tiles_host`, `irlosc_set_tile_kernel`), against the reference's golden outputs and
the oracle.  The lane kernel runs the streaming kernel's per-instance arithmetic (bit-identical); the pair kernel sums
the arms' contributions in another order (equal to rounding, same status flags)."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, GOLDEN_CASES_F4, load_golden
from test_gpu_parity import REL_TOL, _golden_state, _layout_from_dict, _rel_err, _torch

pytestmark = pytest.mark.gpu


TILE_KERNEL_NAME = {"lane": "osc_step_lane", "pair": "osc_step_pair"}


@pytest.mark.parametrize("tile_kernel", ["lane", "pair"])
@pytest.mark.parametrize("packed_M,full6_J", [(False, False), (True, True)])
@pytest.mark.parametrize("case", GOLDEN_CASES + GOLDEN_CASES_F4)
def test_lane_kernel_matches_reference_golden(case, packed_M, full6_J, tile_kernel):
    torch = _torch()
    from irl_control_b200 import _native
    from irl_control_b200.engine import BatchedOSC
    g, ld = load_golden(case)
    layout = _layout_from_dict(ld, topology=True, check=False)
    eng = BatchedOSC(layout, device=0)
    eng.set_tile_kernel(tile_kernel)
    st = _golden_state(g, layout, torch, packed_M, full6_J)
    B = int(st["dq"].shape[0])
    assert eng.tile_entries == len(eng.tile_spec()) > 0
    tiles = eng.pack_tiles(st)
    out = eng.step_tiles(tiles, B, want_u_all=True, target_vel=st.get("target_vel"))
    torch.cuda.synchronize()
    assert eng.last_kernel.startswith(TILE_KERNEL_NAME[tile_kernel]), eng.last_kernel
    ctrl, u_all, status = (out[k].cpu().numpy() for k in ("ctrl", "u_all", "status"))
    bad = g["index_error"]
    assert np.all((status[bad] & _native.ST_DX_RANGE) != 0) and np.all(np.isnan(ctrl[bad])) and np.all(np.isnan(u_all[bad]))
    ok = ~bad
    if ok.any():
        assert np.array_equal((status[ok] & _native.ST_PINV) != 0, g["pinv"][ok])
        assert not np.any(status[ok] & (_native.ST_M_NOT_PD | _native.ST_DX_RANGE | _native.ST_SPARSITY))
        e_u = _rel_err(u_all[ok], g["u_all"][ok])
        e_c = np.abs(ctrl[ok] - g["ctrl"][ok]).max(axis=1) / np.abs(g["u_all"][ok]).max(axis=1)
        print("%s kernel=%s worst rel err u_all %.2e ctrl %.2e" % (case, eng.last_kernel, e_u.max(), e_c.max()))
        assert e_u.max() < REL_TOL and e_c.max() < REL_TOL
        vel = (np.asarray(g["target_vel"]) != 0).all(axis=-1).any(axis=-1)
        assert np.array_equal((status[ok] & _native.ST_VEL_BRANCH) != 0, vel[ok])
    # the host packer writes the same tiles; the host-buffer entry point returns the same results
    host = {k: v.cpu().numpy() for k, v in st.items()}
    th = eng.pack_tiles_host(host)
    assert np.array_equal(th, tiles.cpu().numpy())
    ho = eng.step_tiles_host(th, B, want_u_all=True, target_vel=host.get("target_vel"))
    assert np.array_equal(ho["ctrl"], ctrl, equal_nan=True) and np.array_equal(ho["status"], status)
    # the streaming kernel runs the lane kernel's per-instance code on the same numbers
    eng.set_kernel(9)
    ref = eng.step(st, want_u_all=True)
    eig = _native.ST_EIGEN
    assert torch.equal(ref["status"] & ~eig, out["status"] & ~eig)
    if tile_kernel == "lane":
        assert torch.equal(ref["status"], out["status"])
        assert np.array_equal(ref["u_all"].cpu().numpy(), u_all, equal_nan=True)
    else:
        ru = ref["u_all"].cpu().numpy()
        assert np.array_equal(np.isnan(ru), np.isnan(u_all))
        if ok.any():
            assert _rel_err(u_all[ok], ru[ok]).max() < REL_TOL


@pytest.mark.parametrize("tile_kernel", ["lane", "pair"])
@pytest.mark.parametrize("scenario,B", [("gain_test", 4096), ("admit_test", 8192), ("insertion", 16384), ("worst_case", 65536),
                                        ("iros2022", 4097)])
def test_lane_kernel_matches_oracle_on_baseline_configs(scenario, B, tile_kernel):
    """BASELINE.json configs 2-4 and the k = 13 worst case at their stated batch sizes (ragged last tile included);
    the oracle checks a strided subset, every instance is compared with the streaming kernel bit for bit."""
    torch = _torch()
    from irl_control_b200 import _native
    from irl_control_b200.engine import BatchedOSC
    from irl_control_b200.synthetic import scenario_layout, synth_batch, kernel_inputs, oracle_inputs
    from oracle import osc_numpy
    layout = scenario_layout(scenario)
    st = synth_batch(layout, B, seed=B + 3, device="cuda:0", insertion_schedule=(scenario == "insertion"))
    eng = BatchedOSC(layout, device=0)
    eng.set_tile_kernel(tile_kernel)
    kin = kernel_inputs(st, layout, qM=True)
    tiles = eng.pack_tiles(kin)
    out = eng.step_tiles(tiles, B, want_u_all=True)
    torch.cuda.synchronize()
    assert eng.last_kernel.startswith(TILE_KERNEL_NAME[tile_kernel]), eng.last_kernel
    u_all, status = out["u_all"].cpu().numpy(), out["status"].cpu().numpy()
    assert np.isfinite(u_all).all() and not np.any(status & (_native.ST_M_NOT_PD | _native.ST_DX_RANGE))
    idx = np.arange(0, B, max(1, B // 400))
    ref = osc_numpy.osc_batch(layout.as_dict(), oracle_inputs(st, layout), idx=idx)
    err = _rel_err(u_all[idx], ref["u_all"])
    agree = ((status[idx] & _native.ST_PINV) != 0) == ref["pinv"]
    near = np.abs(np.abs(ref["det"]) - 1e-4) < 1e-9
    assert np.all(agree | near), "branch mismatch away from the det threshold"
    print("%s B=%d kernel=%s: worst rel err %.2e, median %.2e, pinv share %.3f, eigen share %.5f" % (
        scenario, B, eng.last_kernel, err[agree].max(), np.median(err), ref["pinv"].mean(), ((status & _native.ST_EIGEN) != 0).mean()))
    assert err[agree].max() < REL_TOL
    # the thread / pair resolves (nearly) every task-space solve itself: the warp-cooperative eigen-solver is the exception
    assert ((status & _native.ST_EIGEN) != 0).mean() < 2e-4
    eng.set_kernel(9)
    s2 = eng.step(kin, want_u_all=True)
    if tile_kernel == "lane":
        assert torch.equal(s2["u_all"], out["u_all"]) and torch.equal(s2["status"], out["status"]) and torch.equal(s2["ctrl"], out["ctrl"])
    else:
        eig = _native.ST_EIGEN
        assert torch.equal(s2["status"] & ~eig, out["status"] & ~eig)
        assert _rel_err(u_all, s2["u_all"].cpu().numpy()).max() < REL_TOL
    cols = [j for dl in layout.devices for j in dl.actuator_trnids]
    assert torch.equal(out["ctrl"], out["u_all"][:, cols])


@pytest.mark.parametrize("tile_kernel", ["lane", "pair"])
@pytest.mark.parametrize("scenario", ["gain_test", "admit_test", "worst_case"])
def test_rank_deficient_task_rows_go_through_the_warp_eigen_solver(scenario, tile_kernel):
    """Three task rows of one arm zeroed: J M^-1 J^T has three exact zero eigenvalues, more than the thread / pair
    deflates itself, so the instance is handed to the warp-cooperative Jacobi solver (osc_eigen.cuh), which must
    reproduce numpy's pinv(rcond=1e-5) of osc.py:55.  (No reference golden and hardly any random instance reaches
    that path any more.)"""
    torch = _torch()
    from irl_control_b200 import _native
    from irl_control_b200.engine import BatchedOSC
    from irl_control_b200.synthetic import scenario_layout, synth_batch, kernel_inputs, oracle_inputs
    from oracle import osc_numpy
    layout = scenario_layout(scenario)
    B = 333
    st = synth_batch(layout, B, seed=77, device="cuda:0")
    marked = np.zeros(B, dtype=bool)
    marked[::7] = True
    di = next(i for i, d in enumerate(layout.devices) if d.name == "ur5left")
    row0 = sum(int(np.sum(d.ctrlr_dof)) for d in layout.devices[:di])
    sel = torch.from_numpy(marked).to(st["J"].device)
    st["J"][sel, row0:row0 + 3, :] = 0.0                      # the kernels' row-stacked J ...
    for comp in [i for i, on in enumerate(layout.devices[di].ctrlr_dof) if on][:3]:
        st["J6"][sel, di, comp, :] = 0.0                     # ... and the oracle's per-device 6 x n form of it
    eng = BatchedOSC(layout, device=0)
    eng.set_tile_kernel(tile_kernel)
    kin = kernel_inputs(st, layout, qM=True)
    out = eng.step_tiles(eng.pack_tiles(kin), B, want_u_all=True)
    torch.cuda.synchronize()
    u_all, status = out["u_all"].cpu().numpy(), out["status"].cpu().numpy()
    assert np.all((status[marked] & _native.ST_EIGEN) != 0) and np.all((status[marked] & _native.ST_PINV) != 0)
    assert ((status[~marked] & _native.ST_EIGEN) != 0).sum() <= 1
    ref = osc_numpy.osc_batch(layout.as_dict(), oracle_inputs(st, layout))
    assert np.all(ref["pinv"][marked])
    err = _rel_err(u_all, ref["u_all"])
    print("%s %s: worst rel err on the %d rank-deficient instances %.2e, others %.2e" % (
        scenario, eng.last_kernel, marked.sum(), err[marked].max(), err[~marked].max()))
    assert err.max() < REL_TOL


@pytest.mark.parametrize("scenario", ["gain_test", "admit_test"])
def test_pinned_tile_kernel_gives_the_same_bits_at_every_batch_size(scenario):
    """`IRLOSC_TILES_AUTO` serves small and large batches with different kernels (equal to rounding); a caller that
    needs bit-identical results across batch sizes - or across GPU counts of a strong-scaling run - pins one."""
    torch = _torch()
    from irl_control_b200.engine import BatchedOSC
    from irl_control_b200.synthetic import scenario_layout, synth_batch, kernel_inputs
    layout = scenario_layout(scenario)
    eng = BatchedOSC(layout, device=0)
    B = 40000                                   # above one wave of half tiles: auto picks per batch size
    st = synth_batch(layout, B, seed=3, device="cuda:0")
    tiles = eng.pack_tiles(kernel_inputs(st, layout, qM=True))
    small = 4096
    names = {}
    for tile_kernel in ("lane", "pair"):
        eng.set_tile_kernel(tile_kernel)
        big = eng.step_tiles(tiles, B)["ctrl"].clone()
        names[tile_kernel] = eng.last_kernel
        part = eng.step_tiles(tiles[:small // 32].contiguous(), small)["ctrl"]
        assert eng.last_kernel.startswith(TILE_KERNEL_NAME[tile_kernel])
        assert torch.equal(part, big[:small]), tile_kernel
    eng.set_tile_kernel("auto")
    a_big = eng.step_tiles(tiles, B)["ctrl"].clone()
    k_big = eng.last_kernel
    a_small = eng.step_tiles(tiles[:small // 32].contiguous(), small)["ctrl"]
    k_small = eng.last_kernel
    scale = a_big[:small].abs().amax(dim=1, keepdim=True)
    assert ((a_small - a_big[:small]).abs() / scale).max().item() < REL_TOL
    print("%s: auto picks %s at B = %d and %s at B = %d" % (scenario, k_big, B, k_small, small))
    assert k_small.startswith("osc_step_pair")


def test_lane_kernel_small_and_ragged_batches():
    torch = _torch()
    from irl_control_b200.engine import BatchedOSC
    from irl_control_b200.synthetic import scenario_layout, synth_batch, kernel_inputs
    layout = scenario_layout("gain_test")
    eng = BatchedOSC(layout, device=0)
    st = synth_batch(layout, 100, seed=5, device="cuda:0")
    kin = kernel_inputs(st, layout, packed_M=True)
    for tile_kernel in ("lane", "pair", "auto"):
        eng.set_tile_kernel(tile_kernel)
        full = eng.step_tiles(eng.pack_tiles(kin), 100)["ctrl"].clone()
        for B in (1, 15, 16, 17, 31, 32, 33, 64, 99):
            sub = {k: v[:B].contiguous() for k, v in kin.items()}
            out = eng.step_tiles(eng.pack_tiles(sub), B)
            assert out["ctrl"].shape == (B, layout.n_ctrl) and torch.equal(out["ctrl"], full[:B]), (tile_kernel, B)
    assert eng.step_tiles(torch.empty(eng.tiles_shape(0), dtype=torch.float64, device="cuda:0"), 0)["ctrl"].shape[0] == 0
