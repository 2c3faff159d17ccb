"""GPU: time kernel variants (irlosc_set_kernel selector) on one scenario; check they agree."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from irl_control_b200.engine import BatchedOSC
from irl_control_b200.synthetic import scenario_layout, synth_batch, kernel_inputs

sc = sys.argv[1] if len(sys.argv) > 1 else "gain_test"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [2, 3, 4]
packed = (sys.argv[4] != "dense") if len(sys.argv) > 4 else True
L = scenario_layout(sc)
st = synth_batch(L, B, seed=3, device="cuda:0")
kin = kernel_inputs(st, L, packed_M=packed)
eng = BatchedOSC(L, device=0)
ref = None
for v in variants:
    try:
        eng.set_kernel(v)
        out = eng.step(kin, want_u_all=True)
    except Exception as e:
        print("variant %d: %s" % (v, e)); continue
    torch.cuda.synchronize()
    u = out["u_all"].clone()
    if ref is None: ref = u
    diff = ((u - ref).abs().amax(1) / ref.abs().amax(1)).max().item()
    o = {"ctrl": torch.empty(B, L.n_ctrl, dtype=torch.float64, device="cuda:0")}
    for _ in range(5): eng.step(kin, out=o, want_status=False)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
    ev[0].record()
    for i in range(20):
        eng.step(kin, out=o, want_status=False); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(20))
    print("%-50s B=%d  median %.4f ms  min %.4f  -> %.3e steps/s   max rel diff vs first %.1e" % (
        eng.last_kernel, B, ts[10], ts[0], B / (ts[10] * 1e-3), diff))
