"""Quick GPU diagnostic: tiled kernel vs generic kernel vs oracle on every scenario."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from irl_control_b200.engine import BatchedOSC
from irl_control_b200.synthetic import scenario_layout, synth_batch, kernel_inputs, oracle_inputs
from oracle import osc_numpy

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4099
for sc in ("gain_test", "admit_test", "insertion", "worst_case"):
    L = scenario_layout(sc)
    st = synth_batch(L, B, seed=3, device="cuda:0")
    eng = BatchedOSC(L, device=0)
    for packed in (False, True):
        kin = kernel_inputs(st, L, packed_M=packed)
        eng.set_kernel(1)
        a = eng.step(kin, want_u_all=True); torch.cuda.synchronize()
        a = {k: v.clone() for k, v in a.items()}
        eng.set_kernel(0)
        b = eng.step(kin, want_u_all=True); torch.cuda.synchronize()
        name = eng.last_kernel
        ua, ub = a["u_all"].cpu().numpy(), b["u_all"].cpu().numpy()
        sa, sb = a["status"].cpu().numpy(), b["status"].cpu().numpy()
        scale = np.abs(ua).max(axis=1)
        err = np.abs(ua - ub).max(axis=1) / scale
        print("%-10s packed=%d %-44s max rel diff tiled-vs-generic %.2e  nan %d  status diff (pinv bit) %d  eigen share generic %.3f tiled %.3f"
              % (sc, packed, name, np.nanmax(err), int(np.isnan(ub).any(axis=1).sum()), int(((sa ^ sb) & 1).sum()),
                 ((sa & 4) != 0).mean(), ((sb & 4) != 0).mean()))
        worst = np.argsort(-np.nan_to_num(err, nan=1e9))[:3]
        for i in worst:
            print("     inst %d err %.2e status g=%d t=%d" % (i, err[i], sa[i], sb[i]))
    idx = np.arange(0, B, max(1, B // 64))
    ref = osc_numpy.osc_batch(L.as_dict(), oracle_inputs(st, L), idx=idx)
    e = np.abs(ub[idx] - ref["u_all"]).max(axis=1) / np.abs(ref["u_all"]).max(axis=1)
    print("     vs oracle (tiled, packed): max %.2e median %.2e" % (e.max(), np.median(e)))
    # timing
    for kern in (1, 0):
        eng.set_kernel(kern)
        kin = kernel_inputs(st, L, packed_M=True)
        out = {"ctrl": torch.empty(B, L.n_ctrl, dtype=torch.float64, device="cuda:0")}
        for _ in range(3): eng.step(kin, out=out, want_status=False)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): eng.step(kin, out=out, want_status=False)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
        print("     %-44s B=%d %.3f ms  %.3e steps/s" % (eng.last_kernel, B, dt * 1e3, B / dt))
