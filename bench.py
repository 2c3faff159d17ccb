#!/usr/bin/env python3
"""DualUR5 OSC control-steps/sec benchmark (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload gain_test] [--batch 65536]

One "step" of the benchmark = one pass of the fused control-law kernel over a
batch of B synthetic DualUR5 instances per GPU (each instance = one
`OSC.generate`, osc.py:120-210).  `value` counts instances per second over
all GPUs with the state resident in HBM; `e2e` is the same metric through
`irlosc_step_host` (HOST buffers, copies inside the timed region).
`--impl reference` times the CPU restatement of the reference (oracle/) on
all host cores - the reference is pure Python and /root/reference does not
exist on the GPU box, so this is `kind: "port"`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DualUR5 OSC control-steps/sec"
UNIT = "control-steps/s"


# ---------------------------------------------------------------- workload description
def algorithmic_bytes(layout, per_instance_max_vel=True):
    """unique fp64 information per control step.  Symmetric M packed
    (n(n+1)/2), dense J (k x n), dq, bias, EE pose + target pose (7 each per device),
    max_vel (2 per device), F/T frame + raw wrench when admittance; output = packed ctrl."""
    n, k, D = layout.n, layout.k, layout.D
    words = n * (n + 1) // 2 + k * n + n + n + 7 * D + 7 * D
    if per_instance_max_vel:
        words += 2 * D
    if layout.admittance:
        words += 6 * D          # SURVEY counts the rotated wrench (12 for two arms)
    words += layout.n_ctrl
    return 8 * words


def side_legs():
    """tools/side_legs.py in child processes (isolated from this one's CUDA context, bounded by a timeout): paths
    written after the round's GPU budget was spent, each checked before it is timed.  Nothing from here enters
    `value`, `e2e` or `roofline`."""
    tool = os.path.join(ROOT, "tools", "side_legs.py")
    runs = [("default", ["--legs", "qm,qm_admit,iros2022,sequence,coop"], {"IRLOSC_FIXUP_COOP": "0"}, 120),
            ("fixup_coop", ["--legs", "coop"], {"IRLOSC_FIXUP_COOP": "1"}, 60),
            ("qm_tree", ["--legs", "qm_tree,fp64_peak"], {}, 75)]
    out = {}
    for name, extra, env_add, limit in runs:
        env = dict(os.environ, **env_add)
        rec = {"legs": []}
        try:
            proc = subprocess.Popen([sys.executable, tool, "--steps", "10", "--warmup", "3"] + extra, env=env,
                                    stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            try:
                so, se = proc.communicate(timeout=limit)
            except subprocess.TimeoutExpired:
                proc.kill()
                so, se = proc.communicate()
                rec["timeout_s"] = limit
            for ln in so.splitlines():
                try:
                    rec["legs"].append(json.loads(ln))
                except ValueError:
                    pass
            if proc.returncode not in (0, None):
                rec["returncode"] = proc.returncode
                rec["stderr_tail"] = se[-400:]
        except Exception as exc:
            rec["error"] = "%s: %s" % (type(exc).__name__, exc)
        out[name] = rec
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 7] or \
               [r for (_, r) in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(float(r[2]) for r in rows)}


# ---------------------------------------------------------------- CPU baseline (oracle port)
def _cpu_worker(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    layout_dict, batch = args
    from oracle import osc_numpy
    t0 = time.perf_counter()
    osc_numpy.osc_batch(layout_dict, batch)
    return time.perf_counter() - t0


def cpu_baseline(layout, st_host, per_core=192):
    """Times oracle/osc_numpy.py (statement-by-statement numpy port of OSC.generate) on all
    host cores: one process per core, OMP_NUM_THREADS=1, `per_core` instances each."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    total = st_host["M"].shape[0]
    per_core = max(1, min(per_core, total // cores))
    jobs = []
    for c in range(cores):
        sl = slice(c * per_core, (c + 1) * per_core)
        jobs.append((layout.as_dict(), {k: v[sl] for k, v in st_host.items()}))
    os.environ["OMP_NUM_THREADS"] = "1"
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(j[0], {k: v[:2] for k, v in j[1].items()}) for j in jobs])   # warm-up
        t0 = time.perf_counter()
        per = pool.map(_cpu_worker, jobs)
        wall = time.perf_counter() - t0
    n = per_core * cores
    return {"value": n / wall, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d instances of the same workload (%d per core, one process per core, OMP_NUM_THREADS=1), "
                      "oracle/osc_numpy.py; %.0f steps/s/core" % (n, per_core, per_core / (sum(per) / len(per)))}


# ---------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="gain_test")
    ap.add_argument("--batch", type=int, default=65536, help="instances per GPU")
    ap.add_argument("--m-layout", default="packed", choices=["packed", "dense", "qM"],
                    help="packed lower triangle (the SURVEY 8d record, default), dense n x n, or MuJoCo's sparse qM "
                         "(IRLOSC_M_QM: 155 instead of 325 doubles; the roofline still counts the SURVEY record)")
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 generic, 2 tiled")
    ap.add_argument("--sm-margin", type=int, default=-1,
                    help="SMs left free for the gather when --gpus > 1 (-1: 0 for the fused gather, 8 for NCCL)")
    ap.add_argument("--gather-buffers", type=int, default=3, help="gathered output buffers in flight (>= 2)")
    ap.add_argument("--gather", default="fused", choices=["fused", "peer", "nccl"],
                    help="result gather for --gpus > 1: fused = multicast stores if the box has NVLS, else peer stores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fused", action="store_true", help="skip the fused state-provider measurement (SURVEY 8 f1)")
    ap.add_argument("--no-side-legs", action="store_true",
                    help="skip tools/side_legs.py (N = 1 only: checked side measurements of paths without a GPU run of "
                         "their own yet, in a child process; use this flag under ncu)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import numpy as np
    import torch
    from irl_control_b200.synthetic import scenario_layout, synth_batch, kernel_inputs, oracle_inputs

    layout = scenario_layout(args.workload)
    B = args.batch
    config = {"workload": "%s layout (n=%d, k=%d, D=%d, admittance=%s), B=%d per GPU, fp64, M %s, J rows [k][n], "
                          "per-instance max_vel" % (args.workload, layout.n, layout.k, layout.D, layout.admittance,
                                                    B, args.m_layout),
              "batch_per_gpu": B, "l2_policy": "inputs larger than L2 (%.0f MB per step vs 126 MB)" % 0.0}

    # ------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        st = synth_batch(layout, min(B, 8192), seed=0)
        ob = oracle_inputs(st, layout)
        vals = []
        for _ in range(args.warmup + args.steps):
            vals.append(cpu_baseline(layout, ob, per_core=96))
        vals = vals[args.warmup:]
        v = float(np.mean([x["value"] for x in vals]))
        cb = dict(vals[-1], value=v)
        n_inst = 96 * cb["cores"]
        config["l2_policy"] = "n/a (CPU)"
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_inst / v,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "config": config, "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------ B200 arm
    from irl_control_b200.engine import BatchedOSC, pinned_empty
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "8")     # 8 MB gather: a handful of channels saturates it
        dist.init_process_group("nccl", device_id=dev)

    st = synth_batch(layout, B, seed=1000 * rank, device=dev)
    kin = kernel_inputs(st, layout, packed_M=(args.m_layout == "packed"), qM=(args.m_layout == "qM"))
    in_bytes = sum(v.numel() * v.element_size() for v in kin.values())
    config["l2_policy"] = "inputs larger than L2 (%.0f MB read per step vs 126 MB L2)" % (in_bytes / 1e6)
    eng = BatchedOSC(layout, device=local_rank)
    eng.set_kernel(args.kernel)
    # Result gather for N > 1 (the only exchange on this path), NBUF output buffers so that the gather
    # of step i overlaps the kernel of step i+1:
    #   fused : the step kernel stores its ctrl rows straight into every rank's gathered array through
    #           peer-mapped symmetric memory (NVLink stores), followed by a symmetric-memory barrier
    #   nccl  : all_gather_into_tensor on NCCL's stream (fallback, or --gather nccl)
    NBUF = max(2, args.gather_buffers)
    outs = [{"ctrl": torch.empty(B, layout.n_ctrl, dtype=torch.float64, device=dev)} for _ in range(NBUF)]
    out = outs[0]
    gathered = None
    pending = [None] * NBUF
    step_no = [0]
    gather_mode = "none"
    if world > 1:
        gather_mode = "nccl"
        if args.gather in ("fused", "peer"):
            try:
                import torch.distributed._symmetric_memory as symm_mem
                gathered = [symm_mem.empty(world * B, layout.n_ctrl, dtype=torch.float64, device=dev) for _ in range(NBUF)]
                handles = [symm_mem.rendezvous(t, dist.group.WORLD) for t in gathered]
                mc_ptrs = [int(getattr(h, "multicast_ptr", 0) or 0) for h in handles]
                if args.gather == "fused" and all(mc_ptrs):       # one multimem store per row through the switch
                    gather_args = [([], rank * B, mc_ptrs[b]) for b in range(NBUF)]
                    gather_mode = "fused-multicast"
                else:                                             # one store per row per peer
                    gather_args = [([int(h.buffer_ptrs[r]) for r in range(world)], rank * B) for h in handles]
                    gather_mode = "fused-peer"
                side = torch.cuda.Stream(device=dev)
            except Exception as exc:          # no symmetric memory on this box: NCCL gather
                sys.stderr.write("fused gather unavailable (%s), using NCCL\n" % exc)
        if gather_mode == "nccl":
            gathered = [torch.empty(world * B, layout.n_ctrl, dtype=torch.float64, device=dev) for _ in range(NBUF)]

    margin = args.sm_margin if args.sm_margin >= 0 else (8 if gather_mode == "nccl" else 0)
    if world > 1 and margin > 0:
        eng.set_sm_margin(margin)      # the NCCL gather kernel needs a few SMs to overlap the next step's kernel

    def one_step():
        b = step_no[0] % NBUF
        step_no[0] += 1
        if pending[b] is not None:          # the buffer's previous gather must be complete before it is overwritten
            if gather_mode.startswith("fused"):
                torch.cuda.current_stream().wait_event(pending[b])
            else:
                pending[b].wait()
            pending[b] = None
        if gather_mode.startswith("fused"):
            eng.step(kin, out=outs[b], want_status=False, gather=gather_args[b])
            done = torch.cuda.Event()
            done.record()
            with torch.cuda.stream(side):   # cross-GPU barrier off the critical path of the next kernel
                side.wait_event(done)
                handles[b].barrier(channel=b)
                fin = torch.cuda.Event()
                fin.record()
            pending[b] = fin
        else:
            eng.step(kin, out=outs[b], want_status=False)
            if gather_mode == "nccl":
                pending[b] = dist.all_gather_into_tensor(gathered[b], outs[b]["ctrl"], async_op=True)

    def drain():
        for b in range(NBUF):
            if pending[b] is not None:
                if gather_mode.startswith("fused"):
                    torch.cuda.current_stream().wait_event(pending[b])
                else:
                    pending[b].wait()
                pending[b] = None

    for _ in range(max(args.warmup, 3)):
        one_step()
    drain()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()          # all ranks enter the timed region together (after rank 0's sampler start-up)
    launches0 = eng.kernel_launches
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_wall0 = time.time()
    ev[0].record()
    for i in range(args.steps):
        one_step()
        ev[i + 1].record()
    drain()
    ev_end = torch.cuda.Event(enable_timing=True)
    ev_end.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    if world > 1:
        dist.barrier()
    total_ms = ev[0].elapsed_time(ev_end)      # includes the last gather
    launches = eng.kernel_launches - launches0
    # kernel-only time for the roofline: events around the kernel alone, same stream
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in kev:
        a.record()
        eng.step(kin, out=out, want_status=False)
        b.record()
    torch.cuda.synchronize()
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / len(kev)
    kernel_name = eng.last_kernel
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax.item())
    ms_per_step = total_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ------------------------------------------------------------ end to end (host buffers, copies timed)
    e2e = None
    if not args.no_e2e:
        host_in = {}
        for k, v in kin.items():
            buf = pinned_empty(tuple(v.shape))
            buf[...] = v.cpu().numpy()
            host_in[k] = buf
        host_out = {"ctrl": pinned_empty((B, layout.n_ctrl))}
        h2d = sum(a.nbytes for a in host_in.values())
        d2h = host_out["ctrl"].nbytes
        for _ in range(2):
            eng.step_host(host_in, out=host_out, want_status=False)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        reps = max(3, min(args.steps, 10))
        for _ in range(reps):
            eng.step_host(host_in, out=host_out, want_status=False)   # returns after D2H completed
        dt = (time.perf_counter() - t0) / reps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * float(tt.item()),
               "api": "BatchedOSC.step_host -> irlosc_step_host (pinned host buffers, chunked H2D/kernel/D2H pipeline)"}
        assert np.isfinite(host_out["ctrl"]).all()

    # ------------------------------------------------------------ fused state provider
    # Same control steps, but M / J / qfrc_bias / EE poses are computed on the GPU from (q, dq) inside
    # the step kernel (irlosc_step_fused) instead of being read from HBM.  Its inputs (~0.6 KB per
    # instance) fit in L2, so L2 is flushed between the individually timed steps.
    fused = None
    if not args.no_fused:
        from irl_control_b200.synthetic import fused_inputs, scenario_model
        _, model = scenario_model(args.workload)
        eng.set_model(model)
        fin = fused_inputs(st, layout)
        fout = {"ctrl": torch.empty(B, layout.n_ctrl, dtype=torch.float64, device=dev)}
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for _ in range(max(args.warmup, 3)):
            eng.step_fused(fin, out=fout, want_status=False)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l0 = eng.kernel_launches
        fev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in fev:
            flush.zero_()
            a.record()
            eng.step_fused(fin, out=fout, want_status=False)
            b.record()
        torch.cuda.synchronize()
        f_ms = sum(a.elapsed_time(b) for a, b in fev) / len(fev)
        f_launches = eng.kernel_launches - l0
        ft = torch.tensor([f_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ft, op=dist.ReduceOp.MAX)
        f_ms = float(ft.item())
        f_in = sum(v.numel() * v.element_size() for v in fin.values())
        fused = {"value": world * B / (f_ms * 1e-3), "unit": UNIT, "ms_per_step": f_ms, "kernel": eng.last_kernel,
                 "gpu_launches": int(f_launches), "input_bytes_per_step": int(f_in // B),
                 "l2_policy": "256 MB flush write between timed steps (inputs %.0f MB < L2)" % (f_in / 1e6),
                 "api": "BatchedOSC.step_fused -> irlosc_step_fused (q, dq, targets in HBM)"}
        # caller loop fused in: the insertion demo's WP / GRIP state machine per instance
        try:
            from irl_control_b200.sequence import ActionSequence
            # the reference's own 12-entry action list and object offsets (insertion_task.yaml), adapters placed
            # at random per episode (insertion_task.py:341-369), waypoints from set_waypoint_targets (206-268)
            from irl_control_b200 import insertion
            from irl_control_b200.configs import action_config
            acfg = action_config("insertion_task.yaml")
            acts, objs = acfg["insertion_action_sequence"], acfg["nist_action_objects"]
            seq = ActionSequence(layout, acts, active_arm="ur5right")
            ia = seq.active_device
            placed = insertion.random_object_poses(B, "right", objs, rng=np.random.default_rng(7 + rank))
            wp_xyz, wp_quat = insertion.waypoint_poses(acts, objs, placed, st["ee_xyz"][:, ia].cpu().numpy())
            sst = seq.new_state(B, wp_xyz, wp_quat, device=dev)
            sin = {k: v for k, v in fin.items() if k not in ("target_xyz", "target_quat")}
            for _ in range(3):
                eng.step_sequence(sin, seq, sst, out=fout, want_status=False)
            torch.cuda.synchronize()
            sev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            for a, b in sev:
                flush.zero_()
                a.record()
                eng.step_sequence(sin, seq, sst, out=fout, want_status=False)
                b.record()
            torch.cuda.synchronize()
            s_ms = sum(a.elapsed_time(b) for a, b in sev) / len(sev)
            fused["sequence"] = {"value": B / (s_ms * 1e-3), "unit": "episode-steps/s per GPU", "ms_per_step": s_ms,
                                 "api": "BatchedOSC.step_sequence -> irlosc_step_sequence (insertion_task.yaml: 12 actions, "
                                        "randomised adapter poses per episode)"}
        except Exception as exc:      # the sequence step needs two arm devices in the layout
            fused["sequence"] = {"unavailable": str(exc)}
        if not args.no_e2e:
            host_in = {}
            for k, v in fin.items():
                buf = pinned_empty(tuple(v.shape))
                buf[...] = v.cpu().numpy()
                host_in[k] = buf
            host_out = {"ctrl": pinned_empty((B, layout.n_ctrl))}
            for _ in range(2):
                eng.step_fused_host(host_in, out=host_out, want_status=False)
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            reps = max(3, min(args.steps, 10))
            for _ in range(reps):
                eng.step_fused_host(host_in, out=host_out, want_status=False)
            dt = (time.perf_counter() - t0) / reps
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            assert np.isfinite(host_out["ctrl"]).all()
            fused["e2e"] = {"value": world * B / float(tt.item()), "unit": UNIT,
                            "h2d_bytes_per_step": int(sum(a.nbytes for a in host_in.values())),
                            "d2h_bytes_per_step": int(host_out["ctrl"].nbytes), "ms_per_step": 1e3 * float(tt.item()),
                            "api": "BatchedOSC.step_fused_host -> irlosc_step_fused_host (pinned host buffers)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------ roofline + CPU baseline (rank 0)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    abytes = algorithmic_bytes(layout)
    achieved = abytes * B / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        tj = json.load(open(tpath))
        key = "%s_B%d_%s" % (args.workload, B, args.m_layout)
        traffic = tj.get(key, {}).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kernel_name, "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_step": abytes, "input_bytes_per_step": in_bytes // B, "peak_source": peak_src}
    cb = None
    if world == 1 and not args.no_cpu_baseline:
        nsample = (os.cpu_count() or 1) * 192
        sub = {k: v[:nsample] for k, v in st.items()}
        cb = cpu_baseline(layout, oracle_inputs(sub, layout))
    config["gather"] = gather_mode
    side = None
    if world == 1 and not args.no_side_legs:
        side = side_legs()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cb, "fused_state": fused}
    if side is not None:
        line["side_legs"] = side
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
