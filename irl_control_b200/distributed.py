"""Data-parallel sharding of a batch of robot instances over the GPUs of one box.

Robot instances are independent: rank r of G owns the contiguous slice
`shard_range(B, r, G)` of the batch, keeps its inputs resident on its own GPU and runs the
fused step kernel on them.  The only exchange on this path is the gather of the packed
control output `[B_local, n_ctrl]` (osc.py:203-208) - one `all_gather_into_tensor` over NCCL
(NVLink 5 / NVSwitch) on the stream the kernel ran on, or `gloo` on CPU for tests.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of rank's contiguous shard; shards differ by at most one instance."""
    base, extra = divmod(B, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_state(state: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    B = next(iter(state.values())).shape[0]
    a, b = shard_range(B, rank, world)
    return {k: v[a:b].contiguous() for k, v in state.items()}


def gather_ctrl(local_ctrl: torch.Tensor, B: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All ranks get the full `[B, n_ctrl]` output in instance order.

    Equal shards use one `all_gather_into_tensor`; ragged shards are padded to the largest
    shard first (at most one row of padding per rank)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_ctrl = local_ctrl.shape[1]
    sizes = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
    biggest = max(sizes)
    send = local_ctrl
    if send.shape[0] != biggest:
        send = torch.zeros(biggest, n_ctrl, dtype=local_ctrl.dtype, device=local_ctrl.device)
        send[:sizes[rank]] = local_ctrl
    out = torch.empty(world * biggest, n_ctrl, dtype=local_ctrl.dtype, device=local_ctrl.device)
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    if all(s == biggest for s in sizes):
        return out
    return torch.cat([out[r * biggest:r * biggest + sizes[r]] for r in range(world)], dim=0)


class ShardedOSC:
    """One rank's view of a batch sharded over `world` GPUs.

    step_fn(local_state) -> local ctrl tensor; in production that is `BatchedOSC.step(...)["ctrl"]`,
    in CPU tests any stand-in with the same contract."""

    def __init__(self, step_fn: Callable[[Dict[str, torch.Tensor]], torch.Tensor],
                 group: Optional[dist.ProcessGroup] = None):
        self.step_fn = step_fn
        self.group = group

    def step(self, full_state: Dict[str, torch.Tensor]) -> torch.Tensor:
        B = next(iter(full_state.values())).shape[0]
        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        local = shard_state(full_state, rank, world)
        return gather_ctrl(self.step_fn(local), B, self.group)
