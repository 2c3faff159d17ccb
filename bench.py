#!/usr/bin/env python3
"""DualUR5 OSC control-steps/sec benchmark (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload gain_test] [--batch 65536]

One "step" = one launch of a tile kernel (csrc/osc_lane.cuh: a thread per instance, the headline workload; or
csrc/osc_pair.cuh: a lane per arm, 6-row layouts and small batches) over a batch of B synthetic DualUR5 instances per
GPU, each instance one `OSC.generate` (osc.py:120-210).  The state lives in HBM in the package's native batch layout,
batch-interleaved tiles; successive steps read DIFFERENT input sets, so nothing is re-read
from L2.

    value      instances / s over all GPUs, state resident in HBM, result gather (N > 1) inside the timed region
    e2e        the same metric through `BatchedOSC.step_tiles_host` -> `irlosc_step_tiles_host`: tiles in pinned HOST
               memory, H2D + kernel + D2H inside the timed region (`e2e_arrays`: the per-variable-array entry point)
    roofline   SURVEY 8d algorithmic bytes per step x B / kernel time (CUDA events) vs the measured HBM peak, plus the
               same on the bytes the kernel actually moves
    strong     fixed TOTAL batches (65 536 and 262 144) split over the N GPUs
    configs    BASELINE.json configs 2-4, the k = 13 worst case, admit_test and iros2022 at B = 65 536, each checked against
               the oracle and timed; batch_sweep: B = 256 ... 262 144 for gain_test and k = 13 (config 5 at one GPU)
    cpu_baseline / --impl reference : the reference's own `OSC.generate` (unmodified sources staged by
               oracle/stage_ref.py, stub simulator) on all host cores - or the numpy port when the sources are absent
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
N_INPUT_SETS = 3


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


def run_config(workload, layout, B):
    """Identical in both arms (the driver compares it)."""
    tile_b = layout.tile_entries * 8
    return {"workload": "%s layout (n=%d, k=%d, D=%d, admittance=%s), B=%d per GPU, fp64; state in batch-interleaved tiles "
                        "(%d B per instance: tree non-zeros of M and J, dq, bias, poses, targets, max_vel), packed ctrl out"
                        % (workload, layout.n, layout.k, layout.D, layout.admittance, B, tile_b),
            "batch_per_gpu": B,
            "l2_policy": "inputs larger than L2: %d distinct input sets of %.0f MB each, used in rotation (126 MB L2)"
                         % (N_INPUT_SETS, tile_b * B / 1e6)}


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
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
        rows = [r for (t, r) in self.rows if t0 - 0.02 <= t <= t1 + 0.05 and len(r) >= 7] or \
               [r for (_, r) in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(float(r[2]) for r in rows)}


# ---------------------------------------------------------------- CPU baseline (the reference, or its numpy port)
def _cpu_worker(args):
    """One pinned process: `per_core` instances through the reference's OSC.generate (kind "reference") or through
    oracle/osc_numpy.py (kind "port").  Returns seconds of control-law work."""
    core, kind, scenario, layout_dict, batch = args
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        os.sched_setaffinity(0, {core})
    except Exception:
        pass
    if kind == "reference":
        from oracle import ref_harness
        runner = getattr(_cpu_worker, "runner", None)
        if runner is None:
            runner = _cpu_worker.runner = ref_harness.scenario_runner(scenario)
        spent, _calls, _f = ref_harness.time_reference_generate(runner, batch)
        return spent
    from oracle import osc_numpy
    t0 = time.perf_counter()
    osc_numpy.osc_batch(layout_dict, batch)
    return time.perf_counter() - t0


def cpu_kind():
    from oracle import ref_harness
    return "reference" if ref_harness.reference_available() else "port"


def host_cores():
    try:
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return list(range(os.cpu_count() or 1))


def cpu_baseline(scenario, layout, ob, per_core=96, repeats=5, pool=None):
    """Whole-box rate of the CPU path: one pinned single-threaded process per host core, `per_core` instances each,
    wall clock over the slowest worker, median of `repeats` rounds (after one warm-up round)."""
    import multiprocessing as mp
    import numpy as np
    cores = host_cores()
    total = ob["M"].shape[0]
    per_core = max(1, min(per_core, total // len(cores)))
    kind = cpu_kind()
    jobs = []
    for c, core in enumerate(cores):
        sl = slice(c * per_core, (c + 1) * per_core)
        jobs.append((core, kind, scenario, layout.as_dict(), {k: v[sl] for k, v in ob.items()}))
    os.environ["OMP_NUM_THREADS"] = "1"
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(len(cores))
    try:
        pool.map(_cpu_worker, [(j[0], j[1], j[2], j[3], {k: v[:2] for k, v in j[4].items()}) for j in jobs], chunksize=1)
        rates, per = [], []
        for _ in range(repeats):
            t0 = time.perf_counter()
            secs = pool.map(_cpu_worker, jobs, chunksize=1)
            wall = time.perf_counter() - t0
            rates.append(per_core * len(cores) / wall)
            per.append(per_core / (sum(secs) / len(secs)))
    finally:
        if own:
            pool.close()
    src = ("the UNMODIFIED reference OSC.generate (oracle/_ref or /root/reference, stub mujoco_py / transforms3d, fake simulator)"
           if kind == "reference" else "oracle/osc_numpy.py (numpy port; reference sources not staged on this box)")
    return {"value": float(np.median(rates)), "unit": UNIT, "cores": len(cores), "kind": kind,
            "rounds": [round(r, 1) for r in rates], "steps_per_s_per_core": float(np.median(per)),
            "sample": "%d instances of the same workload per round (%d per core, one pinned process per core, "
                      "OMP_NUM_THREADS=1), median of %d rounds; %s" % (per_core * len(cores), per_core, repeats, src)}


# ---------------------------------------------------------------- helpers (GPU arm)
def event_times(torch, fn, n):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for i, (a, b) in enumerate(ev):
        a.record()
        fn(i)
        b.record()
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in ev]


def median(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_for(kernel_name, workload, B):
    """dram__bytes_read + dram__bytes_write of ONE launch from an `ncu --set full` capture (profiles/traffic.json), only
    when it was captured for exactly this kernel instantiation, workload and batch."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.isfile(path):
        return None
    rec = json.load(open(path)).get("%s|%s|B%d" % (kernel_name, workload, B))
    return rec.get("dram_bytes_per_launch") if rec else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="gain_test")
    ap.add_argument("--batch", type=int, default=65536, help="instances per GPU")
    ap.add_argument("--sm-margin", type=int, default=-1,
                    help="SMs left free for the gather when --gpus > 1 (-1: 0 for the fused gather, 8 for NCCL)")
    ap.add_argument("--gather-buffers", type=int, default=3, help="gathered output buffers in flight (>= 2)")
    ap.add_argument("--gather", default="fused", choices=["fused", "peer", "nccl"],
                    help="result gather for --gpus > 1: fused = multicast stores if the box has NVLS, else peer stores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of one CUDA graph of K steps")
    ap.add_argument("--no-extras", action="store_true",
                    help="headline only: skip the strong-scaling, per-config, arrays-layout, fused-state and latency legs "
                         "(use under ncu)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import numpy as np
    import torch
    from irl_control_b200.synthetic import scenario_layout, synth_batch, kernel_inputs, oracle_inputs

    layout = scenario_layout(args.workload)
    B = args.batch
    config = run_config(args.workload, layout, B)

    # ------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        st = synth_batch(layout, 4096, seed=0)
        ob = oracle_inputs(st, layout)
        import multiprocessing as mp
        per_core = 64
        pool = mp.get_context("fork").Pool(len(host_cores()))
        vals = []
        try:
            for _ in range(args.warmup + args.steps):       # one "step" = one round over per_core x cores instances
                vals.append(cpu_baseline(args.workload, layout, ob, per_core=per_core, repeats=1, pool=pool))
        finally:
            pool.close()
        vals = vals[args.warmup:]
        v = float(np.median([x["value"] for x in vals]))
        cb = dict(vals[-1], value=v, rounds=[round(x["value"], 1) for x in vals])
        n_inst = per_core * cb["cores"]
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

    eng = BatchedOSC(layout, device=local_rank)
    E = eng.tile_entries
    assert E == layout.tile_entries and E > 0
    # N_INPUT_SETS distinct batches per rank; tiles are the resident state, the qM arrays feed the arrays-layout legs
    sts = [synth_batch(layout, B, seed=1000 * rank + i, device=dev) for i in range(N_INPUT_SETS)]
    arrays = [kernel_inputs(s, layout, qM=True) for s in sts]
    tiles = [eng.pack_tiles(a) for a in arrays]
    torch.cuda.synchronize()
    tile_bytes = E * 8

    # Result gather for N > 1 (the only exchange on this path), NBUF output buffers so that the gather
    # of step i overlaps the kernel of step i+1:
    #   fused : the step kernel stores its ctrl rows straight into every rank's gathered array through
    #           the NVSwitch multicast mapping (or peer-mapped pointers), followed by a symmetric-memory barrier
    #   nccl  : all_gather_into_tensor on NCCL's stream (fallback, or --gather nccl)
    NBUF = max(2, args.gather_buffers)
    outs = [torch.empty(B, layout.n_ctrl, dtype=torch.float64, device=dev) for _ in range(NBUF)]
    gathered, handles, gather_args, side = None, None, None, None
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
                    gather_args = [([], mc_ptrs[b]) for b in range(NBUF)]
                    gather_mode = "fused-multicast"
                else:                                             # one store per row per peer
                    gather_args = [([int(h.buffer_ptrs[r]) for r in range(world)], 0) for h in handles]
                    gather_mode = "fused-peer"
                side = torch.cuda.Stream(device=dev)
            except Exception as exc:          # no symmetric memory on this box: NCCL gather
                sys.stderr.write("fused gather unavailable (%s), using NCCL\n" % exc)
                gathered = None
        if gather_mode == "nccl":
            gathered = [torch.empty(world * B, layout.n_ctrl, dtype=torch.float64, device=dev) for _ in range(NBUF)]
    margin = args.sm_margin if args.sm_margin >= 0 else (8 if gather_mode == "nccl" else 0)
    if world > 1 and margin > 0:
        eng.set_sm_margin(margin)      # the NCCL gather kernel needs a few SMs to overlap the next step's kernel

    def wait_pending(b):
        if pending[b] is not None:
            if gather_mode.startswith("fused"):
                torch.cuda.current_stream().wait_event(pending[b])
            else:
                pending[b].wait()
            pending[b] = None

    def one_step(_i=None, nloc=B, t_list=tiles):
        """One control step of this rank's shard of `nloc` instances (+ its share of the gather of world x nloc rows)."""
        i = step_no[0]
        b = i % NBUF
        step_no[0] += 1
        wait_pending(b)                     # the buffer's previous gather must be complete before it is overwritten
        o = {"ctrl": outs[b][:nloc]}
        if gather_mode.startswith("fused"):
            ptrs, mc = gather_args[b]
            eng.step_tiles(t_list[i % len(t_list)], nloc, out=o, want_status=False, gather=(ptrs, rank * nloc, mc))
            done = torch.cuda.Event()
            done.record()
            with torch.cuda.stream(side):   # cross-GPU barrier off the critical path of the next kernel
                side.wait_event(done)
                handles[b].barrier(channel=b)
                fin = torch.cuda.Event()
                fin.record()
            pending[b] = fin
        else:
            eng.step_tiles(t_list[i % len(t_list)], nloc, out=o, want_status=False)
            if gather_mode == "nccl":
                pending[b] = dist.all_gather_into_tensor(gathered[b][:world * nloc], outs[b][:nloc], async_op=True)

    def drain():
        for b in range(NBUF):
            wait_pending(b)

    launch_mode = ["eager"]

    def capture(n, **kw):
        """The n steps (kernel launches, side-stream barriers, buffer waits) as ONE CUDA graph: with a ~50 us kernel the
        Python / driver cost of a step (event + barrier + launch calls) would otherwise set the pace."""
        if args.no_graph:
            return None
        try:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            step_no[0] = 0
            with torch.cuda.graph(g):
                for i in range(n):
                    one_step(i, **kw)
                drain()
            torch.cuda.synchronize()
            return g
        except Exception as exc:
            sys.stderr.write("CUDA graph capture unavailable (%s: %s), timing eager launches\n" % (type(exc).__name__, exc))
            for b in range(NBUF):
                pending[b] = None
            torch.cuda.synchronize()
            return None

    region_log = []

    def timed_loop(n, repeat=1, regions=1, **kw):
        """`repeat` x n steps back to back between two events on the launching stream (gathers included); ms per step,
        max over ranks.  regions > 1: that measurement taken `regions` times (each one bracketed by barrier + synchronize
        on both sides), the median is returned and every region is logged - a 1 ms region on several GPUs is at the mercy
        of how far apart the ranks' streams happen to be when it starts."""
        g = capture(n, **kw)
        ok = torch.tensor([1 if g is not None else 0], device=dev)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)      # every rank replays, or none does
        if int(ok.item()) == 0:
            g = None
        launch_mode[0] = "cuda graph of %d steps" % n if g is not None else "eager"
        if g is not None:
            g.replay()                                     # warm-up replay
        times = []
        for _ in range(regions):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
                if gather_mode.startswith("fused"):
                    # a device-side barrier right before the start event: the host barrier leaves the ranks' streams
                    # hundreds of microseconds apart, which the first in-loop barrier would otherwise charge to the region
                    handles[0].barrier(channel=0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(repeat):
                if g is not None:
                    g.replay()
                else:
                    for i in range(n):
                        one_step(i, **kw)
                    drain()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t.item()) / (n * repeat))
        if regions > 1:
            region_log[:] = times
        return sorted(times)[len(times) // 2]

    W = max(args.warmup, 3)
    for i in range(W):
        one_step(i)
    drain()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    t_wall0 = time.time()
    ms_per_step = timed_loop(args.steps, regions=5)
    headline_regions = list(region_log)
    launches = args.steps                  # one lane-kernel launch per step (the N > 1 barrier kernels are torch's)
    headline_launch = launch_mode[0]
    value = world * B / (ms_per_step * 1e-3)
    # the same K steps repeated for >= 1 s: what the rate is once clocks and power have settled
    rep = int(min(40000, max(1, 1.0 / (ms_per_step * 1e-3 * args.steps))))
    sus_ms = timed_loop(args.steps, repeat=rep)
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    sustained = {"steps": rep * args.steps, "ms_per_step": sus_ms, "value": world * B / (sus_ms * 1e-3),
                 "seconds": rep * args.steps * sus_ms * 1e-3}

    # Kernel time for the roofline.  At N = 1 the timed region IS K launches of the kernel back to back, so its average
    # launch duration is the region's time / K.  Events around single eager launches (which also see the ~5 us of
    # launch + event overhead a graph hides) are reported beside it, and are the figure used when the region also
    # contains the gather (N > 1).
    out0 = {"ctrl": outs[0]}
    ks = event_times(torch, lambda i: eng.step_tiles(tiles[i % N_INPUT_SETS], B, out=out0, want_status=False), max(args.steps, 10))
    kernel_ms_isolated = sum(ks) / len(ks)
    kernel_ms = ms_per_step if world == 1 else kernel_ms_isolated
    kernel_name = eng.last_kernel

    # ------------------------------------------------------------ gather verification (N > 1)
    gather_verified = None
    if world > 1:
        drain()
        torch.cuda.synchronize()
        dist.barrier()
        for b in range(NBUF):
            pending[b] = None
        step_no[0] = 0                                   # buffer 0, input set 0 on every rank
        one_step(0)
        drain()
        torch.cuda.synchronize()
        dist.barrier()
        got = gathered[0]
        ref = torch.empty(world * B, layout.n_ctrl, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(ref, outs[0])
        same = bool(torch.equal(got, ref))
        # rank 0 recomputes a foreign shard from scratch (rank 1's input set 0) and compares its gathered rows
        if rank == 0:
            st1 = synth_batch(layout, B, seed=1000 * 1 + 0, device=dev)
            t1 = eng.pack_tiles(kernel_inputs(st1, layout, qM=True))
            o1 = eng.step_tiles(t1, B, want_status=False)
            torch.cuda.synchronize()
            same = same and bool(torch.equal(o1["ctrl"], got[B:2 * B]))
            del st1, t1, o1
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gather_verified = bool(flag.item() == 1)

    # ------------------------------------------------------------ strong scaling: fixed TOTAL batch split over the ranks
    strong = None
    if not args.no_extras:
        strong = []
        for B_total in (65536, 262144):
            nloc = B_total // world
            if nloc * world != B_total or nloc < 32:
                continue
            if nloc > B and world > 1:
                continue
            saved_outs = None
            if nloc <= B:
                # distinct slices of the resident tiles, enough of them that a replay of the K-step graph reads more than
                # the L2 holds (a third of a wave of 8 192 instances is 20 MB: three rotating prefixes would stay in L2)
                nt = (nloc + 31) // 32
                slice_bytes = nt * eng.tile_entries * 32 * 8
                n_off = int(max(1, min((B // 32) // nt, -(-(320 << 20) // (slice_bytes * N_INPUT_SETS)))))
                t_list = [t[j * nt:(j + 1) * nt] for j in range(n_off) for t in tiles]
            else:                                                        # N = 1 at 262 144: a larger resident state
                t_list = []
                for i in range(N_INPUT_SETS):
                    s_big = synth_batch(layout, nloc, seed=5000 + i, device=dev)
                    t_list.append(eng.pack_tiles(kernel_inputs(s_big, layout, qM=True)))
                    del s_big
                saved_outs = list(outs)
                for b in range(NBUF):
                    outs[b] = torch.empty(nloc, layout.n_ctrl, dtype=torch.float64, device=dev)
            for i in range(3):
                one_step(i, nloc=nloc, t_list=t_list)
            drain()
            ms = timed_loop(max(args.steps, 20), nloc=nloc, t_list=t_list)
            strong.append({"total_batch": B_total, "batch_per_gpu": nloc, "ms_per_step": ms, "value": B_total / (ms * 1e-3),
                           "distinct_input_mb": round(len(t_list) * t_list[0].numel() * 8 / 2**20, 1)})
            if saved_outs is not None:
                for b in range(NBUF):
                    outs[b] = saved_outs[b]
            del t_list
        torch.cuda.empty_cache()

    # ------------------------------------------------------------ end to end (host buffers, copies timed)
    e2e, e2e_arrays = None, None
    if not args.no_e2e:
        def host_loop(call):
            for _ in range(2):
                call()
            if world > 1:
                dist.barrier()
            reps = max(3, min(args.steps, 10))
            t0 = time.perf_counter()
            for _ in range(reps):
                call()                                   # returns after the D2H copy completed
            dt = (time.perf_counter() - t0) / reps
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())
        host_out = {"ctrl": pinned_empty((B, layout.n_ctrl))}
        host_tiles = pinned_empty(eng.tiles_shape(B))
        host_tiles[...] = tiles[0].cpu().numpy()
        dt = host_loop(lambda: eng.step_tiles_host(host_tiles, B, out=host_out, want_status=False))
        assert np.array_equal(host_out["ctrl"], eng.step_tiles(tiles[0], B, want_status=False)["ctrl"].cpu().numpy())
        e2e = {"value": world * B / dt, "unit": UNIT, "h2d_bytes_per_step": int(host_tiles.nbytes),
               "d2h_bytes_per_step": int(host_out["ctrl"].nbytes), "ms_per_step": 1e3 * dt,
               "api": "BatchedOSC.step_tiles_host -> irlosc_step_tiles_host (state tiles in pinned host memory, chunked "
                      "H2D / tile kernel / D2H pipeline); a caller that assembles its batch writes tiles directly "
                      "(irlosc_tile_spec), per-variable arrays go through e2e_arrays"}
        del host_tiles
        host_in = {}
        for k, v in arrays[0].items():
            buf = pinned_empty(tuple(v.shape))
            buf[...] = v.cpu().numpy()
            host_in[k] = buf
        dt = host_loop(lambda: eng.step_host(host_in, out=host_out, want_status=False))
        e2e_arrays = {"value": world * B / dt, "unit": UNIT, "h2d_bytes_per_step": int(sum(a.nbytes for a in host_in.values())),
                      "d2h_bytes_per_step": int(host_out["ctrl"].nbytes), "ms_per_step": 1e3 * dt, "kernel": eng.last_kernel,
                      "api": "BatchedOSC.step_host -> irlosc_step_host (MuJoCo-style arrays in pinned host memory: sparse qM, "
                             "J rows, dq, bias, poses, targets)"}
        del host_in

        # What the host side of the box can deliver: every rank copies 256 MB of pinned host memory to its GPU at the same
        # time, nothing else running.  e2e is bounded by this, not by the kernels.
        probe_h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
        probe_d = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        probe_d.copy_(probe_h, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(8):
            probe_d.copy_(probe_h, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        gbs = torch.tensor([8 * probe_h.numel() / (a.elapsed_time(b) * 1e-3) / 1e9], dtype=torch.float64, device=dev)
        lo_, sum_ = gbs.clone(), gbs.clone()
        if world > 1:
            dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
            dist.all_reduce(sum_, op=dist.ReduceOp.SUM)
        h2d_probe = {"h2d_gbs_per_rank_min": float(lo_.item()), "h2d_gbs_aggregate": float(sum_.item()),
                     "e2e_h2d_gbs_aggregate": world * e2e["h2d_bytes_per_step"] / (e2e["ms_per_step"] * 1e-3) / 1e9,
                     "note": "plain pinned-memory cudaMemcpyAsync, all ranks concurrently; the aggregate is the ceiling of e2e"}
        e2e["host_link"] = h2d_probe
        del probe_h, probe_d

    # ------------------------------------------------------------ other legs (every rank runs them; rank 0 reports)
    extras = {}
    if not args.no_extras:
        extras = extra_legs(args, torch, np, eng, layout, sts, arrays, tiles, dev, world, dist, rank)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------ roofline + CPU baseline (rank 0)
    peak, peak_src = hbm_peak()
    abytes = algorithmic_bytes(layout)
    moved = tile_bytes + 8 * layout.n_ctrl
    achieved = abytes * B / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic_for(kernel_name, args.workload, B), "kernel": kernel_name, "kernel_ms": kernel_ms,
                "kernel_ms_source": ("timed region / K (K back-to-back launches, nothing else on the stream)" if world == 1 else
                                     "CUDA events around single eager launches (the timed region also holds the gather)"),
                "kernel_ms_isolated_eager": kernel_ms_isolated,
                "algorithmic_bytes_per_step": abytes, "peak_source": peak_src,
                "moved": {"bytes_per_step": moved, "achieved": moved * B / (kernel_ms * 1e-3) / 1e9,
                          "frac": moved * B / (kernel_ms * 1e-3) / 1e9 / peak,
                          "note": "what one launch reads and writes: the tile (tree non-zeros only) + packed ctrl"}}
    cb = None
    if world == 1 and not args.no_cpu_baseline:
        sub = {k: v[:len(host_cores()) * 96] for k, v in sts[0].items()}
        cb = cpu_baseline(args.workload, layout, oracle_inputs(sub, layout))
    gather_link = None
    if world > 1:
        # the all-gather delivers every other rank's rows into each GPU: NVLink ingress per GPU per step
        bytes_in = (world - 1) * B * layout.n_ctrl * 8
        gather_link = {"bytes_in_per_gpu_per_step": bytes_in, "achieved_gbs": bytes_in / (sustained["ms_per_step"] * 1e-3) / 1e9,
                       "peak_gbs": 770.0, "peak_source": "measured peer-copy bandwidth per direction (B200_PROFILING.md)",
                       "min_ms_per_step_at_peak": bytes_in / 770e9 * 1e3,
                       "note": "when this exceeds the kernel time the step is NVLink-bound, not kernel-bound"}
        gather_link["frac"] = gather_link["achieved_gbs"] / gather_link["peak_gbs"]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cb,
            "sustained": sustained,
            "timed_regions": {"count": len(headline_regions), "ms_per_step": [round(x, 6) for x in headline_regions],
                              "reported": "median; each region is exactly K steps between barrier + synchronize"},
            "launch": headline_launch, "gather": gather_mode, "gather_verified": gather_verified,
            "gather_link": gather_link, "strong": strong,
            "e2e_arrays": e2e_arrays}
    line.update(extras)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def latency_b1(np, median, scenario="gain_test", calls=300):
    """Wall time of one `OSC.generate(targets)` for one robot: this package's Device / Robot / OSC (state pull in Python,
    then irlosc_step_host: H2D, one kernel, D2H, stream sync) and, when its sources are staged, the unmodified
    reference's - same fake simulator class, same loaded instance, same targets."""
    import irl_control_b200 as pkg
    from irl_control_b200 import configs
    from irl_control_b200.dual_ur5 import DualUR5Model
    from irl_control_b200.synthetic import SCENARIOS, oracle_inputs, patched_config, scenario_layout, synth_batch
    from oracle import osc_numpy, ref_harness

    class Sim(ref_harness.FakeSim):
        def full_mass_matrix(self):                      # what `_mj_fullM` yields (robot.py:69-70)
            nv = self.model.nv
            return np.asarray(self.data.qM, dtype=np.float64).reshape(nv, nv)

    sc = SCENARIOS[scenario]
    cfg = patched_config(sc)
    lay = scenario_layout(scenario)
    model = DualUR5Model(n_free_objects=configs.SCENE_FREE_OBJECTS[sc["scene"]])
    sim = Sim(model)
    devices = [pkg.Device(d, model, sim, True) for d in cfg["devices"]]
    robot = pkg.Robot([devices[i] for i in cfg["robots"][0]["device_ids"]], "DualUR5", sim, True)
    by_name = {c["name"]: c for c in cfg["controller_configs"]}
    osc = pkg.OSC(robot, sim, [(dev, dict(by_name[c])) for dev, c in sc["device_cfgs"]], dict(by_name["nullspace"]),
                  admittance=sc["admittance"])
    names = list(sc["targets"])
    st = synth_batch(lay, 64, seed=3)
    ob = oracle_inputs(st, lay)
    inst = {"M": ob["M"][5], "J6": ob["J"][5], "dq": ob["dq"][5], "bias": ob["bias"][5], "ee_xyz": ob["ee_xyz"][5],
            "ee_quat": ob["ee_quat"][5], "ft_xmat": ob["ft_xmat"][5], "ft_raw": ob["ft_raw"][5]}
    sim.load_instance(inst, names, devices)
    targets = {}
    for d, nm in enumerate(names):
        t = pkg.Target(np.zeros(6), np.zeros(6))
        t.set_xyz(ob["tgt_xyz"][5][d])
        t.set_quat(ob["tgt_quat"][5][d])
        targets[nm] = t
    for _ in range(20):
        idxs, forces = osc.generate(targets)
    ref = osc_numpy.osc_batch(lay.as_dict(), {k: v[5:6] for k, v in ob.items()})
    err = float(np.abs(np.concatenate(forces) - ref["ctrl"][0]).max() / np.abs(ref["u_all"][0]).max())
    ts, tg, tc = [], [], []
    for _ in range(calls):
        t0 = time.perf_counter()
        osc.generate(targets)
        ts.append(time.perf_counter() - t0)
    e1 = osc.engine_for(names)
    for _ in range(calls):
        t0 = time.perf_counter()
        st1 = osc.gather_state(targets)
        t1 = time.perf_counter()
        e1.step_host(st1)
        t2 = time.perf_counter()
        tg.append(t1 - t0)
        tc.append(t2 - t1)
    t0 = time.perf_counter()
    osc_numpy.osc_batch(lay.as_dict(), ob)
    lat = {"generate_latency_us": 1e6 * median(ts), "generate_p90_us": 1e6 * sorted(ts)[int(0.9 * len(ts))],
           "state_pull_python_us": 1e6 * median(tg), "c_abi_step_host_us": 1e6 * median(tc),
           "numpy_port_per_call_us": 1e6 * (time.perf_counter() - t0) / 64, "max_rel_err_vs_oracle": err,
           "kernel": e1.last_kernel,
           "note": "B = 1 through Device / Robot / OSC.generate on a fake simulator: state pull (Python) + irlosc_step_host "
                   "(H2D, one kernel, D2H, stream sync); the reference's 1 kHz loop budget is 1000 us"}
    if cpu_kind() == "reference":
        runner = ref_harness.scenario_runner(scenario)
        spent, n_calls, _ = ref_harness.time_reference_generate(runner, {k: v[5:6] for k, v in ob.items()}, repeat=calls)
        lat["reference_generate_per_call_us"] = 1e6 * spent / n_calls
    return lat


def extra_legs(args, torch, np, eng, layout, sts, arrays, tiles, dev, world, dist, rank):
    """Measurements beside the headline: the per-variable-array kernels, the tile packer, the fused state provider
    with its caller loop, BASELINE.json configs 2-4 + the worst case (checked against the oracle, N = 1 only), and the
    B = 1 latency of the drop-in `OSC.generate`."""
    from irl_control_b200.engine import BatchedOSC, pinned_empty
    from irl_control_b200.synthetic import scenario_layout, synth_batch, kernel_inputs, oracle_inputs
    B = args.batch
    peak, _ = hbm_peak()
    out = {"ctrl": torch.empty(B, layout.n_ctrl, dtype=torch.float64, device=dev)}
    n = max(args.steps, 10)
    res = {}

    def allmax(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- per-variable arrays (MuJoCo's qM + J rows + ...): the record-staging tree kernel / streaming kernel, and the packer
    eng.set_kernel(0)
    for i in range(3):
        eng.step(arrays[i % N_INPUT_SETS], out=out, want_status=False)
    ms = allmax(median(event_times(torch, lambda i: eng.step(arrays[i % N_INPUT_SETS], out=out, want_status=False), n)))
    in_b = sum(v.numel() * v.element_size() for v in arrays[0].values()) // B
    res["arrays_layout"] = {"value": world * B / (ms * 1e-3), "ms_per_step": ms, "kernel": eng.last_kernel,
                            "input_bytes_per_step": int(in_b), "roofline_frac": algorithmic_bytes(layout) * B / (ms * 1e-3) / 1e9 / peak,
                            "api": "BatchedOSC.step -> irlosc_step (qM, J rows, dq, bias, poses, targets as separate arrays in HBM)"}
    tb = torch.empty_like(tiles[0])
    for i in range(3):
        eng.pack_tiles(arrays[i % N_INPUT_SETS], tiles=tb)
    ms = allmax(median(event_times(torch, lambda i: eng.pack_tiles(arrays[i % N_INPUT_SETS], tiles=tb), n)))
    res["arrays_layout"]["pack_to_tiles_ms"] = ms
    del tb

    # ---- fused state provider and the caller loop inside the step (f2): q, dq in, torques out
    try:
        from irl_control_b200.synthetic import fused_inputs, scenario_model
        _, model = scenario_model(args.workload)
        eng.set_model(model)
        fins = [fused_inputs(s, layout) for s in sts]
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for i in range(3):
            eng.step_fused(fins[i % N_INPUT_SETS], out=out, want_status=False)
        ts = []
        for i in range(n):
            flush.zero_()
            ts += event_times(torch, lambda _i: eng.step_fused(fins[i % N_INPUT_SETS], out=out, want_status=False), 1)
        ms = allmax(median(ts))
        f_in = sum(v.numel() * v.element_size() for v in fins[0].values())
        res["fused_state"] = {"value": world * B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "kernel": eng.last_kernel,
                              "input_bytes_per_step": int(f_in // B),
                              "l2_policy": "256 MB flush write between timed steps (inputs %.0f MB < L2)" % (f_in / 1e6),
                              "parity": "against the package's own rigid-body model (dual_ur5.py); MuJoCo parity unpinned",
                              "api": "BatchedOSC.step_fused -> irlosc_step_fused (q, dq, targets in HBM)"}
        try:
            from irl_control_b200.sequence import ActionSequence
            from irl_control_b200 import insertion
            from irl_control_b200.configs import action_config
            acfg = action_config("insertion_task.yaml")
            acts, objs = acfg["insertion_action_sequence"], acfg["nist_action_objects"]
            seq = ActionSequence(layout, acts, active_arm="ur5right")
            ia = seq.active_device
            placed = insertion.random_object_poses(B, "right", objs, rng=np.random.default_rng(7 + rank))
            wp_xyz, wp_quat = insertion.waypoint_poses(acts, objs, placed, sts[0]["ee_xyz"][:, ia].cpu().numpy())
            sst = seq.new_state(B, wp_xyz, wp_quat, device=dev)
            sin = {k: v for k, v in fins[0].items() if k not in ("target_xyz", "target_quat")}
            for _ in range(3):
                eng.step_sequence(sin, seq, sst, out=out, want_status=False)
            ts = []
            for i in range(n):
                flush.zero_()
                ts += event_times(torch, lambda _i: eng.step_sequence(sin, seq, sst, out=out, want_status=False), 1)
            s_ms = allmax(median(ts))
            res["fused_state"]["sequence"] = {"value": world * B / (s_ms * 1e-3), "unit": "episode-steps/s", "ms_per_step": s_ms,
                                              "api": "BatchedOSC.step_sequence -> irlosc_step_sequence (insertion_task.yaml: 12 "
                                                     "actions, randomised adapter poses per episode)"}
        except Exception as exc:      # the sequence step needs two arm devices in the layout
            res["fused_state"]["sequence"] = {"unavailable": str(exc)}
        if not args.no_e2e:
            host_in = {}
            for k, v in fins[0].items():
                buf = pinned_empty(tuple(v.shape))
                buf[...] = v.cpu().numpy()
                host_in[k] = buf
            host_out = {"ctrl": pinned_empty((B, layout.n_ctrl))}
            for _ in range(2):
                eng.step_fused_host(host_in, out=host_out, want_status=False)
            if world > 1:
                dist.barrier()
            reps = max(3, min(args.steps, 10))
            t0 = time.perf_counter()
            for _ in range(reps):
                eng.step_fused_host(host_in, out=host_out, want_status=False)
            dt = allmax((time.perf_counter() - t0) / reps)
            res["fused_state"]["e2e"] = {"value": world * B / dt, "unit": UNIT,
                                         "h2d_bytes_per_step": int(sum(a.nbytes for a in host_in.values())),
                                         "d2h_bytes_per_step": int(host_out["ctrl"].nbytes), "ms_per_step": 1e3 * dt,
                                         "api": "BatchedOSC.step_fused_host -> irlosc_step_fused_host (pinned host buffers)"}
        del flush
    except Exception as exc:
        res["fused_state"] = {"unavailable": "%s: %s" % (type(exc).__name__, exc)}

    if world > 1 or rank != 0:
        return res

    # ---- BASELINE.json configs 2-4 and the worst case: oracle check on a strided subset, then the kernel time
    from oracle import osc_numpy
    def time_config(wl, Bc, check=True):
        lay = scenario_layout(wl)
        en = BatchedOSC(lay, device=dev.index)
        ss = [synth_batch(lay, Bc, seed=77 + i, device=dev, insertion_schedule=(wl == "insertion")) for i in range(N_INPUT_SETS)]
        tl = [en.pack_tiles(kernel_inputs(s, lay, qM=True)) for s in ss]
        o = en.step_tiles(tl[0], Bc, want_u_all=True)
        torch.cuda.synchronize()
        st_ = o["status"].cpu().numpy()
        r = {"workload": wl, "B": Bc, "k": lay.k}
        if check:
            idx = np.arange(0, Bc, max(1, Bc // 96))
            ref = osc_numpy.osc_batch(lay.as_dict(), oracle_inputs(ss[0], lay), idx=idx)
            got = o["u_all"].cpu().numpy()[idx]
            r["max_rel_err_vs_oracle"] = float((np.abs(got - ref["u_all"]).max(axis=1) / np.abs(ref["u_all"]).max(axis=1)).max())
            r["oracle_instances"] = int(len(idx))
        oc = {"ctrl": torch.empty(Bc, lay.n_ctrl, dtype=torch.float64, device=dev)}
        ts = []
        if tl[0].numel() * 8 * N_INPUT_SETS < (300 << 20):           # small inputs: flush L2 between steps instead
            fl = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
            for i in range(n):
                fl.zero_()
                ts += event_times(torch, lambda _i: en.step_tiles(tl[i % N_INPUT_SETS], Bc, out=oc, want_status=False), 1)
            policy = "256 MB flush write between timed steps"
            del fl
        else:
            ts = event_times(torch, lambda i: en.step_tiles(tl[i % N_INPUT_SETS], Bc, out=oc, want_status=False), n)
            policy = "inputs larger than L2 (rotating sets)"
        ms = median(ts)
        ab = algorithmic_bytes(lay)
        mv = en.tile_entries * 8 + 8 * lay.n_ctrl
        r.update({"kernel": en.last_kernel, "ms": ms, "value": Bc / (ms * 1e-3),
                  "roofline": {"frac": ab * Bc / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_step": ab,
                               "moved_bytes_per_step": mv, "moved_frac": mv * Bc / (ms * 1e-3) / 1e9 / peak},
                  "pinv_share": float(((st_ & 1) != 0).mean()), "warp_finished_share": float(((st_ & 4) != 0).mean()),
                  "l2_policy": policy})
        en.close()
        del ss, tl, o, oc
        torch.cuda.empty_cache()
        return r

    res["configs"] = [time_config(wl, Bc) for wl, Bc in (("gain_test", 4096), ("admit_test", 8192), ("insertion", 16384),
                                                         ("worst_case", 65536), ("admit_test", 65536), ("iros2022", 65536))]
    # SURVEY 8d config 5 at one GPU: the batch sweep of the headline layout and of the k = 13 worst case (each step is
    # timed by its own events; the N > 1 points of config 5 are the weak headline and `strong`)
    res["batch_sweep"] = {wl: [{k_: v for k_, v in time_config(wl, Bs, check=False).items()
                                if k_ in ("B", "kernel", "ms", "value", "roofline", "l2_policy")}
                               for Bs in (256, 1024, 4096, 16384, 65536, 262144)] for wl in ("gain_test", "worst_case")}


    # ---- B = 1: the drop-in OSC.generate (examples/gain_test.py:143-147 calls it every 1 ms) next to the reference's own,
    #      both on the same stand-in simulator (oracle.ref_harness.FakeSim, cheap array accessors) and the same state
    try:
        res["latency_b1"] = latency_b1(np, median)
    except Exception as exc:
        res["latency_b1"] = {"unavailable": "%s: %s" % (type(exc).__name__, exc)}
    return res


if __name__ == "__main__":
    try:
        main()
    except BaseException:                      # never leave a half-dead process behind on the GPU box
        import traceback
        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)
