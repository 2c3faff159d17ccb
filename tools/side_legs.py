#!/usr/bin/env python3
"""Side measurements of paths that have not had a GPU run of their own yet.

`bench.py` (N = 1 only) runs this script in a CHILD process with a timeout and attaches what it
prints to its JSON line under `"side_legs"`; nothing here feeds `value`, `e2e` or `roofline`.  A child
process, because these paths were written after round 1's GPU budget was spent: if one of them faults,
the parent's measurement and its JSON line are unaffected.  Every leg prints one JSON line as soon as it is
done; each is first CHECKED (against the validated packed-`M` path or the oracle) and then timed with CUDA
events on the launching stream, warm-up first, inputs larger than L2 or an L2 flush in between.

Legs (`--legs a,b,...`):
  qm        IRLOSC_M_QM (MuJoCo's sparse qM as the `M` input): bit-equality with the packed-M run of the
            streaming kernel, kernel time of both, and the host-buffer (pinned, H2D + D2H timed) rate
  qm_admit  the same kernel comparison on the admit_test layout (k = 12)
  iros2022  SURVEY 8 (f4) layout: parity against the oracle on a strided subset, kernel times
  sequence  the reference's 12-entry insertion action list, randomised adapters, B = 16 384
  qm_tree   the tree-sparse kernel's qM instantiations (explicit kernel selection; never run on a GPU before - bench.py
            gives this leg a child process of its own)
  fp64_peak cuBLAS DGEMM rate of the box (the FP64 reference point SURVEY 8d asks for)
  coop      fused gain_test step with IRLOSC_FIXUP_COOP=1 (run by bench.py as a second child with that
            environment variable): parity against the oracle + step time
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DEV = "cuda:0"          # tests/test_side_legs_mock.py runs the legs' host logic with "cpu" and a stand-in engine


def emit(name, **kw):
    print(json.dumps(dict(leg=name, **kw)), flush=True)


def time_steps(torch, fn, steps, warmup, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        if flush is not None:
            flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return {"ms_mean": sum(ts) / len(ts), "ms_median": ts[len(ts) // 2], "ms_min": ts[0]}


def rel_err(np, got, want):
    scale = np.abs(want).max(axis=1, keepdims=True)
    return float((np.abs(got - want) / scale).max())


def leg_qm(torch, np, scenario, B, steps, warmup, name, with_e2e):
    from irl_control_b200.engine import BatchedOSC, pinned_empty
    from irl_control_b200.synthetic import kernel_inputs, scenario_layout, synth_batch
    layout = scenario_layout(scenario)
    st = synth_batch(layout, B, seed=4242, device=DEV)
    eng = BatchedOSC(layout, device=0)
    eng.set_kernel(9)
    pin, qin = kernel_inputs(st, layout, packed_M=True), kernel_inputs(st, layout, qM=True)
    a = eng.step(pin, want_u_all=True)
    b = eng.step(qin, want_u_all=True)
    torch.cuda.synchronize()
    same = bool(torch.equal(a["u_all"], b["u_all"]) and torch.equal(a["ctrl"], b["ctrl"]) and torch.equal(a["status"], b["status"]))
    res = {"workload": scenario, "batch": B, "bit_identical_to_packed": same, "kernel": eng.last_kernel,
           "input_bytes_per_step": {"packed": int(sum(v.numel() * 8 for v in pin.values()) // B),
                                    "qM": int(sum(v.numel() * 8 for v in qin.values()) // B)}}
    if not same:
        res["max_rel_diff"] = rel_err(np, b["u_all"].cpu().numpy(), a["u_all"].cpu().numpy())
        emit(name, **res)
        return
    out = {"ctrl": torch.empty_like(a["ctrl"])}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    res["packed"] = time_steps(torch, lambda: eng.step(pin, out=out, want_status=False), steps, warmup, flush)
    res["qM"] = time_steps(torch, lambda: eng.step(qin, out=out, want_status=False), steps, warmup, flush)
    res["l2_policy"] = "256 MB flush write before every timed step"
    res["steps_per_s"] = {k: B / (res[k]["ms_median"] * 1e-3) for k in ("packed", "qM")}
    if with_e2e:
        # launch knobs the library reads per launch (experiments only): warps per CTA and the cp.async L2 hint
        sweep = {}
        for lay, inp in (("packed", pin), ("qM", qin)):
            for warps in (6, 7, 8):
                for mode in (0, 1):
                    os.environ["IRLOSC_STREAM_WARPS"], os.environ["IRLOSC_STREAM_MODE"] = str(warps), str(mode)
                    t = time_steps(torch, lambda: eng.step(inp, out=out, want_status=False), max(3, steps // 2), 2, flush)
                    sweep["%s_w%d_mode%d" % (lay, warps, mode)] = t["ms_median"]
        os.environ.pop("IRLOSC_STREAM_WARPS", None)
        os.environ.pop("IRLOSC_STREAM_MODE", None)
        res["stream_knob_sweep_ms"] = sweep
    if with_e2e:
        host_in = {}
        for k, v in qin.items():
            buf = pinned_empty(tuple(v.shape))
            buf[...] = v.cpu().numpy()
            host_in[k] = buf
        host_out = {"ctrl": pinned_empty((B, layout.n_ctrl))}
        eng.set_kernel(0)
        for _ in range(2):
            eng.step_host(host_in, out=host_out, want_status=False)
        ok = bool(np.array_equal(host_out["ctrl"], a["ctrl"].cpu().numpy()))
        reps = 8
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.step_host(host_in, out=host_out, want_status=False)
        dt = (time.perf_counter() - t0) / reps
        res["e2e_qM"] = {"value": B / dt, "unit": "control-steps/s", "ms_per_step": 1e3 * dt, "equal_to_device_run": ok,
                         "h2d_bytes_per_step": int(sum(x.nbytes for x in host_in.values())),
                         "d2h_bytes_per_step": int(host_out["ctrl"].nbytes),
                         "api": "BatchedOSC.step_host({'qM': ...}) -> irlosc_step_host, IRLOSC_M_QM"}
    emit(name, **res)


def leg_qm_tree(torch, np, B, steps, warmup):
    """Tree-sparse kernel staging qM (osc_step_tree<..., qM>, explicit selection only): first run on a GPU.  Equality
    with the packed-M tree kernel on a full and on a ragged batch, then the variants' kernel times."""
    from irl_control_b200.engine import BatchedOSC
    from irl_control_b200.synthetic import kernel_inputs, scenario_layout, synth_batch
    layout = scenario_layout("gain_test")
    eng = BatchedOSC(layout, device=0)
    eng.set_kernel(2)
    res = {"workload": "gain_test", "batch": B}
    for nb in (1003, B):
        st = synth_batch(layout, nb, seed=99, device=DEV)
        pin, qin = kernel_inputs(st, layout, packed_M=True), kernel_inputs(st, layout, qM=True)
        a = eng.step(pin, want_u_all=True)
        name_p = eng.last_kernel
        b = eng.step(qin, want_u_all=True)
        torch.cuda.synchronize()
        same = bool(torch.equal(a["u_all"], b["u_all"]) and torch.equal(a["ctrl"], b["ctrl"]) and torch.equal(a["status"], b["status"]))
        res["bit_identical_to_packed_B%d" % nb] = same
        if not same:
            res["max_rel_diff_B%d" % nb] = rel_err(np, b["u_all"].cpu().numpy(), a["u_all"].cpu().numpy())
            emit("qm_tree", **res)
            return
    out = {"ctrl": torch.empty_like(a["ctrl"])}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    res[name_p] = time_steps(torch, lambda: eng.step(pin, out=out, want_status=False), steps, warmup, flush)
    for which in (2, 5, 6):                                   # qM table variants 0 (w8 s4, else w7 s3), 3 (w7 s4), 4 (w8 s3)
        try:
            eng.set_kernel(which)
            eng.step(qin, out=out, want_status=False)
            nm = eng.last_kernel
            res[nm] = time_steps(torch, lambda: eng.step(qin, out=out, want_status=False), steps, warmup, flush)
        except Exception as exc:
            res["selector_%d" % which] = "%s: %s" % (type(exc).__name__, exc)
    emit("qm_tree", **res)


def leg_iros2022(torch, np, B, steps, warmup):
    from irl_control_b200 import _native
    from irl_control_b200.engine import BatchedOSC
    from irl_control_b200.synthetic import fused_inputs, kernel_inputs, oracle_inputs, scenario_model, synth_batch
    from oracle import osc_numpy
    layout, model = scenario_model("iros2022")
    st = synth_batch(layout, B, seed=31, device=DEV)
    eng = BatchedOSC(layout, device=0)
    kin = kernel_inputs(st, layout, packed_M=True)
    o = eng.step(kin, want_u_all=True)
    torch.cuda.synchronize()
    name = eng.last_kernel
    idx = np.arange(0, B, max(1, B // 256))
    ref = osc_numpy.osc_batch(layout.as_dict(), oracle_inputs(st, layout), idx=idx)
    status = o["status"].cpu().numpy()
    agree = ((status[idx] & _native.ST_PINV) != 0) == ref["pinv"]
    res = {"workload": "iros2022 (k=13, targets base, ur5left, ur5right)", "batch": B, "kernel": name,
           "max_rel_err_vs_oracle": rel_err(np, o["u_all"].cpu().numpy()[idx][agree], ref["u_all"][agree]),
           "branch_agreement": float(agree.mean()), "pinv_share": float(ref["pinv"].mean()), "checked": int(len(idx))}
    eng.set_model(model)
    fin = fused_inputs(st, layout)
    f = eng.step_fused(fin, want_u_all=True)
    torch.cuda.synchronize()
    res["fused_max_rel_err_vs_oracle"] = rel_err(np, f["u_all"].cpu().numpy()[idx][agree], ref["u_all"][agree])
    res["fused_kernel"] = eng.last_kernel
    out = {"ctrl": torch.empty_like(o["ctrl"])}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    res["step"] = time_steps(torch, lambda: eng.step(kin, out=out, want_status=False), steps, warmup, flush)
    res["step_fused"] = time_steps(torch, lambda: eng.step_fused(fin, out=out, want_status=False), steps, warmup, flush)
    emit("iros2022", **res)


def leg_sequence(torch, np, B, steps, warmup):
    from irl_control_b200 import insertion
    from irl_control_b200.configs import action_config
    from irl_control_b200.engine import BatchedOSC
    from irl_control_b200.sequence import ActionSequence
    from irl_control_b200.synthetic import fused_inputs, scenario_model, synth_batch
    layout, model = scenario_model("insertion")
    st = synth_batch(layout, B, seed=77, device=DEV, insertion_schedule=True)
    eng = BatchedOSC(layout, device=0)
    eng.set_model(model)
    cfg = action_config("insertion_task.yaml")
    acts, objs = cfg["insertion_action_sequence"], cfg["nist_action_objects"]
    seq = ActionSequence(layout, acts, active_arm="ur5right")
    ia = seq.active_device
    placed = insertion.random_object_poses(B, "right", objs, rng=np.random.default_rng(11))
    wp_xyz, wp_quat = insertion.waypoint_poses(acts, objs, placed, st["ee_xyz"][:, ia].cpu().numpy())
    sst = seq.new_state(B, wp_xyz, wp_quat, device=DEV)
    fin = fused_inputs(st, layout)
    sin = {k: v for k, v in fin.items() if k not in ("target_xyz", "target_quat")}
    out = {"ctrl": torch.empty(B, layout.n_ctrl, dtype=torch.float64, device=DEV)}
    eng.step_sequence(sin, seq, sst, out=out)
    torch.cuda.synchronize()
    # first step: every episode is in action 0 and its active-arm target is the first waypoint
    ok = bool((sst["action"] == 0).all().item()) and bool(torch.equal(sst["target_xyz"][:, ia].cpu(), torch.from_numpy(wp_xyz[:, 0])))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    t = time_steps(torch, lambda: eng.step_sequence(sin, seq, sst, out=out, want_status=False), steps, warmup, flush)
    emit("sequence", workload="insertion layout, insertion_task.yaml (12 actions), randomised adapters", batch=B,
         first_step_state_ok=ok, finite=bool(torch.isfinite(out["ctrl"]).all().item()), step=t,
         episode_steps_per_s=B / (t["ms_median"] * 1e-3))


def leg_fp64_peak(torch, np):
    """FP64 rate of the box (not in MEASURED_PEAKS.json): cuBLAS DGEMM 8192^3 through torch, best of 5 -
    the denominator for the FP64-pipe share the fused kernel is quoted against."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=DEV)
    b = torch.randn(n, n, dtype=torch.float64, device=DEV)
    c = torch.empty_like(a)
    t = time_steps(torch, lambda: torch.matmul(a, b, out=c), 5, 2)
    emit("fp64_peak", how="torch.matmul float64 %d^3 (cuBLAS DGEMM), CUDA events, best of 5" % n,
         tflops=2.0 * n ** 3 / (t["ms_min"] * 1e-3) / 1e12, ms_min=t["ms_min"])


def leg_coop(torch, np, B, steps, warmup):
    from irl_control_b200.engine import BatchedOSC
    from irl_control_b200.synthetic import fused_inputs, oracle_inputs, scenario_model, synth_batch
    from oracle import osc_numpy
    res = {"IRLOSC_FIXUP_COOP": os.environ.get("IRLOSC_FIXUP_COOP", "0"), "batch": B}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for scenario in ("gain_test", "admit_test"):
        layout, model = scenario_model(scenario)
        st = synth_batch(layout, B, seed=5, device=DEV)
        eng = BatchedOSC(layout, device=0)
        eng.set_model(model)
        fin = fused_inputs(st, layout)
        o = eng.step_fused(fin, want_u_all=True)
        torch.cuda.synchronize()
        status = o["status"].cpu().numpy()
        hard = np.nonzero(status & 0x04)[0][:192]                     # IRLOSC_ST_EIGEN: went through the fix-up
        idx = np.unique(np.concatenate([np.arange(0, B, max(1, B // 128)), hard]))
        ref = osc_numpy.osc_batch(layout.as_dict(), oracle_inputs(st, layout), idx=idx)
        agree = ((status[idx] & 1) != 0) == ref["pinv"]
        out = {"ctrl": torch.empty_like(o["ctrl"])}
        res[scenario] = {"max_rel_err_vs_oracle": rel_err(np, o["u_all"].cpu().numpy()[idx][agree], ref["u_all"][agree]),
                         "branch_agreement": float(agree.mean()), "fixups": int((status & 0x04 != 0).sum()),
                         "fixups_checked": int(len(hard)),
                         "step_fused": time_steps(torch, lambda: eng.step_fused(fin, out=out, want_status=False), steps, warmup, flush)}
    emit("coop", **res)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--legs", default="qm,qm_admit,iros2022,sequence")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    import numpy as np
    import torch
    assert torch.cuda.is_available(), "needs a GPU"
    torch.cuda.set_device(0)
    table = {
        "qm": lambda: leg_qm(torch, np, "gain_test", args.batch, args.steps, args.warmup, "qm", True),
        "qm_admit": lambda: leg_qm(torch, np, "admit_test", args.batch, args.steps, args.warmup, "qm_admit", False),
        "iros2022": lambda: leg_iros2022(torch, np, args.batch, args.steps, args.warmup),
        "sequence": lambda: leg_sequence(torch, np, min(args.batch, 16384), args.steps, args.warmup),
        "coop": lambda: leg_coop(torch, np, args.batch, args.steps, args.warmup),
        "qm_tree": lambda: leg_qm_tree(torch, np, args.batch, args.steps, args.warmup),
        "fp64_peak": lambda: leg_fp64_peak(torch, np),
    }
    for leg in args.legs.split(","):
        try:
            table[leg]()
        except Exception as exc:            # a failed leg is reported, the others still run
            emit(leg, error="%s: %s" % (type(exc).__name__, exc))


if __name__ == "__main__":
    main()
