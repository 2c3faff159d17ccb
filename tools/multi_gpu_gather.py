"""2+ GPUs: the fused gather of irlosc_step / irlosc_step_tiles (peer stores, and NVSwitch multicast stores when the
box offers them) equals an NCCL all_gather of the local outputs.  Run under torchrun (tests/test_gpu_multi.py does)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    from irl_control_b200.engine import BatchedOSC
    from irl_control_b200.synthetic import scenario_layout, synth_batch, kernel_inputs
    ok = True
    for scenario, B, kern, mode in (("gain_test", 4099, 0, "peer"), ("admit_test", 1024, 0, "peer"), ("gain_test", 777, 1, "peer"),
                                    ("gain_test", 4096, 0, "peer"), ("gain_test", 4096, 0, "multicast"),
                                    ("gain_test", 4099, 0, "multicast"), ("admit_test", 1031, 2, "multicast"),
                                    # streaming kernel (kernel 9; auto for 6-row arm devices), incl. warp-finished instances
                                    ("gain_test", 4099, 9, "peer"), ("gain_test", 4096, 9, "multicast"),
                                    ("worst_case", 2051, 9, "multicast"), ("worst_case", 2048, 0, "peer"),
                                    ("admit_test", 4096, 0, "multicast"),
                                    # lane kernel on tiles
                                    ("gain_test", 4096, "lane", "multicast"), ("gain_test", 4099, "lane", "peer"),
                                    ("worst_case", 8195, "lane", "multicast"), ("admit_test", 2048, "lane", "peer")):
        layout = scenario_layout(scenario)
        st = synth_batch(layout, B, seed=100 + rank, device=dev)
        kin = kernel_inputs(st, layout, packed_M=True)
        eng = BatchedOSC(layout, device=lr)
        if kern != "lane":
            eng.set_kernel(kern)
        g = symm_mem.empty(world * B, layout.n_ctrl, dtype=torch.float64, device=dev)
        g.fill_(float("nan"))
        h = symm_mem.rendezvous(g, dist.group.WORLD)
        mc = int(getattr(h, "multicast_ptr", 0) or 0)
        if mode == "multicast" and mc == 0:
            if rank == 0:
                print(scenario, B, "multicast mapping not available on this box: skipped")
            continue
        h.barrier(channel=0)
        gather = ([int(p) for p in h.buffer_ptrs], rank * B) if mode == "peer" else ([], rank * B, mc)
        if kern == "lane":
            out = eng.step_tiles(eng.pack_tiles(kin), B, want_status=False, gather=gather)
        else:
            out = eng.step(kin, want_status=False, gather=gather)
        h.barrier(channel=0)
        torch.cuda.synchronize()
        ref = torch.empty(world * B, layout.n_ctrl, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(ref, out["ctrl"])
        same = torch.equal(g, ref)
        ok = ok and same
        if rank == 0:
            print(scenario, B, eng.last_kernel, mode, "fused gather == nccl all_gather:", same)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("ALL OK" if flag.item() == 1 else "MISMATCH")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
