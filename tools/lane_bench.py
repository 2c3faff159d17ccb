#!/usr/bin/env python3
"""GPU experiment driver (not part of the product): times the step kernels of one scenario at one batch size and
checks that they agree.  Inputs rotate over several distinct buffers so that no step re-reads what the previous one
left in L2.

    python tools/lane_bench.py --scenario gain_test --batch 65536 --threads 256,384,448 --prefetch 0,2
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timeit(torch, fn, iters, warmup=3):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i, (a, b) in enumerate(ev):
        a.record()
        fn(i)
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenario", default="gain_test")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--threads", default="256,320,384,448,512")
    ap.add_argument("--prefetch", default="0,2")
    ap.add_argument("--staged", default="1", help="comma-separated: 1 = TMA-staged ring, 0 = direct LDG")
    ap.add_argument("--stages", default="3", help="comma-separated ring depths for the staged form")
    ap.add_argument("--nbuf", type=int, default=3)
    ap.add_argument("--sweep-pair", type=int, default=0, help="CTA size of the pair kernel for the --sweep runs (0: default launch)")
    ap.add_argument("--pair", default="", help="comma-separated CTA sizes of the pair kernel (two lanes per instance)")
    ap.add_argument("--others", default="tree_packed,tree_qm,stream_qm,fused,pack")
    ap.add_argument("--sweep", default="", help="comma-separated batch sizes: time the default lane launch on prefixes of the tiles")
    args = ap.parse_args()
    import torch
    from irl_control_b200.engine import BatchedOSC
    from irl_control_b200.synthetic import scenario_layout, synth_batch, kernel_inputs, fused_inputs, scenario_model
    dev = torch.device("cuda", 0)
    layout = scenario_layout(args.scenario)
    B = args.batch
    eng = BatchedOSC(layout, device=0)
    sts = [synth_batch(layout, B, seed=11 + i, device=dev) for i in range(args.nbuf)]
    kq = [kernel_inputs(s, layout, qM=True) for s in sts]
    kp = [kernel_inputs(s, layout, packed_M=True) for s in sts]
    tiles = [eng.pack_tiles(k) for k in kq]
    out = {"ctrl": torch.empty(B, layout.n_ctrl, dtype=torch.float64, device=dev)}
    E = eng.tile_entries
    res = []

    def rec(name, med, mn, **kw):
        r = dict(name=name, scenario=args.scenario, B=B, ms=round(med, 5), ms_min=round(mn, 5), kernel=eng.last_kernel,
                 steps_per_s=B / (med * 1e-3), **kw)
        res.append(r)
        print(json.dumps(r), flush=True)

    eng.set_kernel(9)
    ref = eng.step(kq[0], want_status=True)
    ref = {k: v.clone() for k, v in ref.items()}
    for th in [int(x) for x in args.threads.split(",") if x]:
        for stg in [int(x) for x in args.staged.split(",") if x]:
            variants = [("stages", int(x)) for x in args.stages.split(",") if x] if stg else \
                       [("prefetch", int(x)) for x in args.prefetch.split(",") if x]
            for key, val in variants:
                os.environ["IRLOSC_LANE_THREADS"] = str(th)
                os.environ["IRLOSC_LANE_STAGED"] = str(stg)
                os.environ["IRLOSC_LANE_STAGES" if stg else "IRLOSC_LANE_PREFETCH"] = str(val)
                o = eng.step_tiles(tiles[0], B, want_status=True)
                same = bool(torch.equal(o["ctrl"], ref["ctrl"]) and torch.equal(o["status"], ref["status"]))
                med, mn = timeit(torch, lambda i: eng.step_tiles(tiles[i % args.nbuf], B, out=out, want_status=False), args.iters)
                rec("lane", med, mn, threads=th, staged=stg, **{key: val}, equal_to_stream=same, tile_bytes_per_instance=E * 8)
    for k_ in ("IRLOSC_LANE_THREADS", "IRLOSC_LANE_PREFETCH", "IRLOSC_LANE_STAGED", "IRLOSC_LANE_STAGES"):
        os.environ.pop(k_, None)
    for th in [int(x) for x in args.pair.split(",") if x]:
        for pf in [int(x) for x in args.prefetch.split(",") if x]:
            os.environ["IRLOSC_PAIR"] = str(th)
            os.environ["IRLOSC_LANE_PREFETCH"] = str(pf)
            o = eng.step_tiles(tiles[0], B, want_status=True)
            fin = torch.isfinite(ref["ctrl"])
            err = float(((o["ctrl"] - ref["ctrl"]).abs()[fin].max() / ref["ctrl"].abs()[fin].max()).item())
            same = bool(torch.equal(o["status"], ref["status"]) and torch.equal(torch.isfinite(o["ctrl"]), fin))
            med, mn = timeit(torch, lambda i: eng.step_tiles(tiles[i % args.nbuf], B, out=out, want_status=False), args.iters)
            rec("pair", med, mn, threads=th, prefetch=pf, status_equal=same, max_rel_err_vs_stream=err)
    os.environ.pop("IRLOSC_PAIR", None)
    os.environ.pop("IRLOSC_LANE_PREFETCH", None)
    if args.sweep_pair:
        os.environ["IRLOSC_PAIR"] = str(args.sweep_pair)
    for Bs in [int(x) for x in args.sweep.split(",") if x]:
        nt = (Bs + 31) // 32
        sub = [t[:nt] for t in tiles]
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        ts = []
        for i in range(args.iters + 3):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eng.step_tiles(sub[i % args.nbuf], Bs, out={"ctrl": out["ctrl"][:Bs]}, want_status=False); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts = sorted(ts[3:])
        r = dict(name="lane_sweep", scenario=args.scenario, B=Bs, ms=round(ts[len(ts) // 2], 5), ms_min=round(ts[0], 5),
                 kernel=eng.last_kernel, steps_per_s=Bs / (ts[len(ts) // 2] * 1e-3), l2="flushed between steps")
        print(json.dumps(r), flush=True)
    os.environ.pop("IRLOSC_PAIR", None)
    others = [x for x in args.others.split(",") if x]
    scale = ref["ctrl"].abs().amax(dim=1, keepdim=True)
    if "tree_packed" in others:
        eng.set_kernel(0)
        o = eng.step(kp[0], want_status=True)
        err = ((o["ctrl"] - ref["ctrl"]).abs() / scale).max().item()
        med, mn = timeit(torch, lambda i: eng.step(kp[i % args.nbuf], out=out, want_status=False), args.iters)
        rec("auto_packed", med, mn, max_rel_diff_vs_stream=err)
    if "tree_qm" in others:
        eng.set_kernel(0)
        o = eng.step(kq[0], want_status=True)
        err = ((o["ctrl"] - ref["ctrl"]).abs() / scale).max().item()
        med, mn = timeit(torch, lambda i: eng.step(kq[i % args.nbuf], out=out, want_status=False), args.iters)
        rec("auto_qM", med, mn, max_rel_diff_vs_stream=err)
    if "stream_qm" in others:
        eng.set_kernel(9)
        med, mn = timeit(torch, lambda i: eng.step(kq[i % args.nbuf], out=out, want_status=False), args.iters)
        rec("stream_qM", med, mn)
    eng.set_kernel(0)
    if "pack" in others:
        tb = torch.empty_like(tiles[0])
        med, mn = timeit(torch, lambda i: eng.pack_tiles(kq[i % args.nbuf], tiles=tb), args.iters)
        rec("pack_qM_to_tiles", med, mn)
    if "fused" in others:
        _, model = scenario_model(args.scenario)
        eng.set_model(model)
        fin = [fused_inputs(s, layout) for s in sts]
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        o = eng.step_fused(fin[0], want_status=True)
        err = ((o["ctrl"] - ref["ctrl"]).abs() / scale).max().item()

        def f(i):
            eng.step_fused(fin[i % args.nbuf], out=out, want_status=False)
        # inputs are small (fit in L2): flush between steps, timed around the step only
        for i in range(3):
            f(i)
        ev = []
        for i in range(args.iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); f(i); b.record()
            ev.append((a, b))
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in ev)
        rec("fused", ts[len(ts) // 2], ts[0], max_rel_diff_vs_stream=err)


if __name__ == "__main__":
    main()
