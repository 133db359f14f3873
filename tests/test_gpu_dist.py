"""GPU, >= 2 devices: launches tests/dist_gpu_check.py under torchrun (NCCL) — sharded ITC + data-parallel heads against the
single-GPU plan on the concatenated batch.  Skipped on single-GPU boxes (the gloo test covers the collective logic)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("extra", [["--mode", "peer"], ["--mode", "peer", "--graph"], ["--mode", "peer", "--graph", "--live"], ["--mode", "nccl"]])
def test_dist_head_plan_two_gpus(extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29517 + len(extra)), os.path.join(ROOT, "tests", "dist_gpu_check.py")] + extra
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    # on failure: the ranks' own lines (checks, tracebacks, device-side "tic:" messages) rather than the launcher's boilerplate
    keep = [ln for ln in r.stdout.splitlines() if not any(t in ln for t in ("elastic", "site-packages/torch/distributed", "OMP_NUM_THREADS", "*****"))]
    marks = [ln for ln in keep if "dist check" in ln or "tic:" in ln]
    assert r.returncode == 0, "\n".join([ln[:160] for ln in marks] + keep[-45:])[-7000:]
