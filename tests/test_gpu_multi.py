"""GPU (>= 2 devices): the fused peer-store gather of `irlosc_step` equals an NCCL all_gather.
Skipped on single-GPU boxes; `tools/multi_gpu_gather.py` is the 2-rank program."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_fused_gather_matches_nccl_all_gather():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tools", "multi_gpu_gather.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "ALL OK" in out.stdout, out.stdout[-2000:]
