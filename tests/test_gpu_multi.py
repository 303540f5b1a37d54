"""-m gpu: 2-rank NCCL run of the sharded A-FAN step vs the single-process global-batch step
(skipped on a 1-GPU box; `gpurun --gpus 2` runs it)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
@pytest.mark.parametrize("graph", [False, True])
def test_two_rank_sharded_step_matches_global_batch(graph, exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(29516 + int(graph) + 2 * (exchange == "nccl")),
           os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    if graph:
        cmd.append("--graph")
    if exchange == "nccl":
        cmd.append("--nccl")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
