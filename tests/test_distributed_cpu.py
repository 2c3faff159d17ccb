"""CPU: the multi-GPU path (shard -> step -> gather) with world_size 2 over gloo.
The per-shard compute is stood in by the oracle (no GPU here); what is tested is that the
gathered result equals the single-process result, for equal and ragged shards."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, B, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from irl_control_b200.distributed import ShardedOSC
    from irl_control_b200.synthetic import scenario_layout, synth_batch, oracle_inputs
    from oracle import osc_numpy
    layout = scenario_layout("gain_test")
    st = synth_batch(layout, B, seed=5)

    def step_fn(local):
        ob = oracle_inputs(local, layout)
        return torch.from_numpy(osc_numpy.osc_batch(layout.as_dict(), ob)["ctrl"])

    out = ShardedOSC(step_fn).step(st)
    if rank == 0:
        ref = osc_numpy.osc_batch(layout.as_dict(), oracle_inputs(st, layout))["ctrl"]
        q.put(bool(np.array_equal(out.numpy(), ref)))
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7])
def test_two_rank_gather_equals_single_process(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + B) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_shard_ranges_cover_the_batch():
    from irl_control_b200.distributed import shard_range
    for B in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 4, 8):
            spans = [shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
