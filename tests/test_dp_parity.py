"""Data-parallel parity on the GPU box: the DataParallel wrapper's averaged gradients equal the single-GPU
gradients of the concatenated batch, including under gradient accumulation and zero_grad(set_to_none=False).

Two transports: `gloo` with both ranks on cuda:0 (runs on the 1-GPU box the driver uses: same bucketing / hook / budget code,
the collective goes through the host) and `nccl` with one GPU per rank (skipped when fewer than 2 GPUs are visible)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(backend, depth, timeout, dp_mode="nccl"):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_worker.py"), backend, str(depth), dp_mode]
    env = dict(os.environ, NCCL_MAX_CTAS="8", OMP_NUM_THREADS="4")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.count("dp_worker ok") == 2, r.stdout[-3000:] + r.stderr[-3000:]


def test_dp_parity_gloo_two_ranks_one_gpu():
    _run("gloo", 3, 600)


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dp_parity_nccl_two_gpus():
    _run("nccl", 12, 900)


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dp_parity_nvls_fused_step_two_gpus():
    """the optimizer step fused with its collectives over NVSwitch multicast (csrc/dp_nvls.cu): parameters after clipped Adam
    steps (with gradient accumulation) equal the single-GPU run on the concatenated batch; ranks bit-identical."""
    _run("nccl", 4, 900, dp_mode="nvls")
