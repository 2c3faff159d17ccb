"""CPU: the host logic of tools/side_legs.py (the checked side measurements bench.py attaches at N = 1) with a
stand-in engine that runs the host builds of the kernels - so that a typo in a leg does not cost the first GPU
numbers of the paths it measures.  Timing values are meaningless here; only structure and the checks are asserted."""
import importlib.util
import json
import os
import time

import numpy as np
import pytest

import fused_host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Event:
    def __init__(self, enable_timing=True):
        self.t = 0.0

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return 1e3 * (other.t - self.t) + 1e-6


class _Engine:
    """`BatchedOSC` stand-in on the host builds (tests/host_fused)."""

    def __init__(self, layout, device=None):
        self.layout, self.model, self.last_kernel = layout, None, "host build"

    def set_kernel(self, which):
        pass

    def set_model(self, model):
        self.model = model

    @staticmethod
    def _np(state):
        return {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in state.items()}

    def _finish(self, res, out, torch_out):
        import torch
        conv = (lambda a: torch.from_numpy(a)) if torch_out else (lambda a: a)
        full = {k: conv(res[k]) for k in ("ctrl", "u_all", "status")}
        if out is not None:
            for k in out:
                out[k][...] = full[k]
        return full

    def step(self, state, out=None, want_u_all=False, want_status=True):
        st = self._np(state)
        if st.get("ft_xmat") is not None and st["ft_xmat"].ndim == 4:
            st["ft_xmat"] = st["ft_xmat"].reshape(st["ft_xmat"].shape[0], -1, 9)
        return self._finish(fused_host.run_stream(self.layout, st), out, True)

    def step_host(self, state, out=None, want_u_all=False, want_status=True):
        return self._finish(fused_host.run_stream(self.layout, dict(state)), out, False)

    def step_fused(self, state, out=None, want_u_all=False, want_status=True, want_ee=False):
        return self._finish(fused_host.run(self.layout, self.model, self._np(state)), out, True)

    def step_sequence(self, state, seq, seq_state, out=None, want_u_all=False, want_status=True):
        res = fused_host.sequence_step(self.layout, self.model, seq, self._np(state), self._np(seq_state))
        return self._finish(res, out, True)


@pytest.fixture()
def legs(monkeypatch):
    import torch
    import irl_control_b200.engine as engine
    spec = importlib.util.spec_from_file_location("side_legs", os.path.join(ROOT, "tools", "side_legs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(mod, "DEV", "cpu")
    monkeypatch.setattr(engine, "BatchedOSC", _Engine)
    monkeypatch.setattr(engine, "pinned_empty", lambda shape, dtype=np.float64: np.empty(shape, dtype=dtype))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    return mod


def _run(capsys, fn):
    fn()
    lines = [json.loads(l) for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and "error" not in lines[0], lines
    return lines[0]


def test_qm_legs(legs, capsys):
    import torch
    r = _run(capsys, lambda: legs.leg_qm(torch, np, "gain_test", 96, 2, 1, "qm", True))
    assert r["bit_identical_to_packed"] and r["input_bytes_per_step"] == {"packed": 4784, "qM": 3424}
    assert r["e2e_qM"]["equal_to_device_run"] and r["e2e_qM"]["h2d_bytes_per_step"] == 96 * 3424
    assert set(r["steps_per_s"]) == {"packed", "qM"}
    r = _run(capsys, lambda: legs.leg_qm(torch, np, "admit_test", 64, 2, 1, "qm_admit", False))
    assert r["bit_identical_to_packed"] and "e2e_qM" not in r


def test_iros2022_sequence_and_coop_legs(legs, capsys):
    import torch
    r = _run(capsys, lambda: legs.leg_iros2022(torch, np, 128, 2, 1))
    assert r["max_rel_err_vs_oracle"] < 1e-6 and r["fused_max_rel_err_vs_oracle"] < 1e-6 and r["branch_agreement"] == 1.0
    r = _run(capsys, lambda: legs.leg_sequence(torch, np, 64, 2, 1))
    assert r["first_step_state_ok"] and r["finite"]
    r = _run(capsys, lambda: legs.leg_coop(torch, np, 128, 2, 1))
    for sc in ("gain_test", "admit_test"):
        assert r[sc]["max_rel_err_vs_oracle"] < 1e-6 and r[sc]["branch_agreement"] == 1.0


def test_qm_tree_leg(legs, capsys):
    import torch
    r = _run(capsys, lambda: legs.leg_qm_tree(torch, np, 64, 2, 1))
    assert r["bit_identical_to_packed_B1003"] and r["bit_identical_to_packed_B64"] and "host build" in r
