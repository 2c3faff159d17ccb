"""2-GPU probe: NCCL all_gather latency/bandwidth for the ctrl gather, and P2P copy bandwidth."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
for n in (65536 * 15, 65536 * 15 * 8):
    x = torch.randn(n, dtype=torch.float64, device=dev); out = torch.empty(world * n, dtype=torch.float64, device=dev)
    for _ in range(5): dist.all_gather_into_tensor(out, x)
    torch.cuda.synchronize(); dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(20): dist.all_gather_into_tensor(out, x)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 20
    if rank == 0: print("all_gather %d B per rank: %.3f ms  (%.1f GB/s per rank)" % (n * 8, ms, n * 8 / ms / 1e6))
if rank == 0:
    print("can_device_access_peer(0,1):", torch.cuda.can_device_access_peer(0, 1))
    a = torch.empty(1 << 28, dtype=torch.uint8, device="cuda:0"); b = torch.empty(1 << 28, dtype=torch.uint8, device="cuda:1")
    for _ in range(3): b.copy_(a)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): b.copy_(a)
    torch.cuda.synchronize(); print("peer copy 256 MiB: %.1f GB/s" % (10 * (1 << 28) / (time.perf_counter() - t0) / 1e9))
dist.destroy_process_group()
