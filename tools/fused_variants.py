#!/usr/bin/env python3
"""A/B timing of the fused-step kernel variants (irlosc_set_kernel(h, 2 + v)) on one GPU.

    python tools/fused_variants.py [--workload gain_test] [--batch 65536] [--variants 0,1,2,3,4]

CUDA events around each step, 256 MB L2 flush between steps (the inputs fit in L2)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="gain_test")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--variants", default="0,1,2,3,4")
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    import torch
    from irl_control_b200 import _native
    from irl_control_b200.engine import BatchedOSC
    from irl_control_b200.synthetic import fused_inputs, scenario_model, synth_batch
    layout, model = scenario_model(args.workload)
    st = synth_batch(layout, args.batch, seed=0, device="cuda:0")
    fin = fused_inputs(st, layout)
    eng = BatchedOSC(layout, device=0)
    eng.set_model(model)
    out = {"ctrl": torch.empty(args.batch, layout.n_ctrl, dtype=torch.float64, device="cuda:0")}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    ref = None
    for v in [int(x) for x in args.variants.split(",")]:
        eng.set_kernel(0 if v == 0 else 2 + v)
        try:
            for _ in range(3):
                eng.step_fused(fin, out=out, want_status=False)
        except _native.OscError as exc:
            print(json.dumps({"variant": v, "error": str(exc)}))
            continue
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in ev:
            flush.zero_()
            a.record()
            eng.step_fused(fin, out=out, want_status=False)
            b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)
        ctrl = out["ctrl"].clone()
        if ref is None:
            ref = ctrl
        print(json.dumps({"variant": v, "kernel": eng.last_kernel, "workload": args.workload, "batch": args.batch,
                          "ms_median": ms[len(ms) // 2], "ms_min": ms[0],
                          "steps_per_s": args.batch / (ms[len(ms) // 2] * 1e-3),
                          "max_abs_diff_vs_first": float((ctrl - ref).abs().max())}))


if __name__ == "__main__":
    main()
