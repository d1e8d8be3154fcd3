"""GPU, >= 2 devices: the sharded alignment loop (images split over ranks, gradient records exchanged inside
align_loop_kernel through peer-mapped memory) under torchrun + NCCL -- tools/check_sharded_align.py asserts that the
sharded result is bit-identical on every rank, agrees with the single-GPU run to fp32 rounding and survives repeated
calls.  40 images / 4 windows; with 8 devices the partition has ranks owning 4-7 images (multi-rank test
at G >= 8 ranks).  Skipped on a single-GPU box (the round-end GPU test box has one)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
def test_sharded_alignment_under_torchrun():
    n = min(8, torch.cuda.device_count())
    n = 1 << (n.bit_length() - 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(REPO, "tools", "check_sharded_align.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    assert r.returncode == 0 and "SHARDED_ALIGN_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
