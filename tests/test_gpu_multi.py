"""GPU, >= 2 devices: launches tests/multigpu_check.py with torchrun (one process per GPU, NCCL) — N-rank gradients ==
mean of the single-rank gradients, replicas bit-identical after graphed training steps, train() under data parallelism.
Skipped on a single-GPU box."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_data_parallel_path_on_two_gpus():
    script = Path(__file__).resolve().parent / "multigpu_check.py"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", str(script)]
    env = dict(os.environ)
    env.pop("B200SEG_DETERMINISTIC", None)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "MULTIGPU_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
