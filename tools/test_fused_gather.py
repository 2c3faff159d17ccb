"""2+ GPUs: the fused peer-store gather of irlosc_step equals an NCCL all_gather of the local outputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
from irl_control_b200.engine import BatchedOSC
from irl_control_b200.synthetic import scenario_layout, synth_batch, kernel_inputs
ok = True
for scenario, B, kern in (("gain_test", 4099, 0), ("admit_test", 1024, 0), ("gain_test", 777, 1)):
    layout = scenario_layout(scenario)
    st = synth_batch(layout, B, seed=100 + rank, device=dev)
    kin = kernel_inputs(st, layout, packed_M=True)
    eng = BatchedOSC(layout, device=lr); eng.set_kernel(kern)
    g = symm_mem.empty(world * B, layout.n_ctrl, dtype=torch.float64, device=dev)
    g.fill_(float("nan"))
    h = symm_mem.rendezvous(g, dist.group.WORLD)
    h.barrier(channel=0)
    out = eng.step(kin, want_status=False, gather=([int(p) for p in h.buffer_ptrs], rank * B))
    h.barrier(channel=0)
    torch.cuda.synchronize()
    ref = torch.empty(world * B, layout.n_ctrl, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(ref, out["ctrl"])
    same = torch.equal(g, ref)
    ok = ok and same
    if rank == 0: print(scenario, B, eng.last_kernel, "fused gather == nccl all_gather:", same)
flag = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0: print("ALL OK" if flag.item() == 1 else "MISMATCH")
dist.destroy_process_group()
